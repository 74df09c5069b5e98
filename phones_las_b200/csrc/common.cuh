// Shared helpers for the phones-las B200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

namespace plas {

// ---- error plumbing (thread-local message behind plas_last_error) -------------------
char* err_buf();
int set_err(int code, const char* fmt, ...);

#define PLAS_OK 0
#define PLAS_EINVAL (-1)
#define PLAS_ECUDA (-2)
#define PLAS_EUNSUPPORTED (-3)

#define PLAS_CUDA(expr)                                                                     \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return plas::set_err(PLAS_ECUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,         \
                           cudaGetErrorString(_e));                                         \
  } while (0)

#define PLAS_REQUIRE(cond, ...)                                                             \
  do {                                                                                      \
    if (!(cond)) return plas::set_err(PLAS_EINVAL, __VA_ARGS__);                            \
  } while (0)

int num_sms();

// ---- device math ---------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// LSTM gate math of tf.nn.rnn_cell.LSTMCell (forget_bias = 1.0 added at run time).
__device__ __forceinline__ void lstm_gates(float zi, float zj, float zf, float zo, float c_prev,
                                           float& c, float& h) {
  c = sigmoidf_acc(zf + 1.0f) * c_prev + sigmoidf_acc(zi) * tanhf(zj);
  h = sigmoidf_acc(zo) * tanhf(c);
}

// Cheaper gate math for the bf16 kernels, still ~2e-7 relative: sigmoid through ex2.approx/rcp.approx
// (no cancellation: 1 + e^-x >= 1) and tanh as the Cephes odd polynomial below 0.625 (where
// 1 - 2/(e^2x + 1) would lose relative accuracy to cancellation) and the exponential form above.
// e^x as ONE MUFU.EX2: __expf without -ftz wraps ex2.approx in a denormal-result path (compare, halve, square: +3 instructions
// per call); results that small vanish in the 1 + e^x that follows, everything else is bit-identical.
__device__ __forceinline__ float exp_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
  return y;
}
__device__ __forceinline__ float sigmoidf_fast(float x) { return __fdividef(1.0f, 1.0f + exp_ftz(-x)); }
__device__ __forceinline__ float tanhf_fast(float x) {
  const float ax = fabsf(x);
  const float x2 = x * x;
  float p = fmaf(x2, -5.70498872745e-3f, 2.06390887954e-2f);
  p = fmaf(p, x2, -5.37397155531e-2f);
  p = fmaf(p, x2, 1.33314422036e-1f);
  p = fmaf(p, x2, -3.33332819422e-1f);
  const float small = fmaf(x * x2, p, x);
  const float big = copysignf(1.0f - __fdividef(2.0f, exp_ftz(2.0f * ax) + 1.0f), x);
  return ax < 0.625f ? small : big;
}
__device__ __forceinline__ void lstm_gates_fast(float zi, float zj, float zf, float zo, float c_prev, float& c,
                                                float& h) {
  c = sigmoidf_fast(zf + 1.0f) * c_prev + sigmoidf_fast(zi) * tanhf_fast(zj);
  h = sigmoidf_fast(zo) * tanhf_fast(c);
}

__device__ __forceinline__ float bf16_round(float x) {
  return __bfloat162float(__float2bfloat16_rn(x));
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) {
  return __bfloat162float(v);
}
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

// Counter-based dropout mask (training; tf.nn.rnn_cell.DropoutWrapper(input_keep_prob), las/ops.py:14-18): element `idx` of
// the tensor identified by `seed` is kept when the top 24 bits of a murmur3-finalised hash fall below keep_prob.  No mask
// is stored: the backward pass regenerates it from the same (seed, idx).  Mirrored in numpy by train.dropout_mask.
__host__ __device__ __forceinline__ uint32_t drop_hash(uint64_t idx, uint32_t seed) {
  uint32_t x = (uint32_t)idx * 0x9E3779B1u + (uint32_t)(idx >> 32) * 0x85EBCA77u + seed;
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  x += seed * 0x27D4EB2Fu;
  x ^= x >> 15; x *= 0x2C1B3C6Du; x ^= x >> 12; x *= 0x297A2D39u; x ^= x >> 15;
  return x;
}
// seed of a tensor at optimiser step s = seed(step 0) + s * DROP_STEP_MUL: the kernels read s from device memory, so a
// captured CUDA graph draws fresh masks on every replay
constexpr uint32_t DROP_STEP_MUL = 0x85EBCA77u;
// multiplier applied to a dropped-out input: 1/keep when kept, 0 otherwise (thresh = keep_prob * 2^24)
__device__ __forceinline__ float drop_scale(uint64_t idx, uint32_t seed, uint32_t thresh, float inv_keep) {
  return (drop_hash(idx, seed) >> 8) < thresh ? inv_keep : 0.f;
}

// standard normal from the counter-based hash (Box-Muller on two 24-bit uniforms); numpy mirrors: train.reference_noise,
// train.reference_weight_noise
__device__ __forceinline__ float hash_normal(uint64_t idx, uint32_t seed) {
  const float u1 = ((float)(drop_hash(idx, seed) >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u2 = ((float)(drop_hash(idx, seed + 1u) >> 8) + 0.5f) * (1.0f / 16777216.0f);
  return sqrtf(-2.0f * logf(u1)) * cosf(6.283185307179586f * u2);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// order-preserving float <-> uint mapping (for atomicMax on floats); 0 sorts below everything.
__device__ __forceinline__ unsigned float_to_ordered(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// ---- inter-CTA signalling through L2 --------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// one elected lane of a fully converged warp (lets the compiler keep descriptor operands in uniform
// registers instead of emitting a per-thread uniformisation loop around tcgen05 / TMA instructions)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---- shared-memory addresses and mbarriers ---------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  uint32_t spins = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && ++spins > (1u << 26)) __trap();  // a protocol bug must fail loudly, not hang the GPU
  } while (!done);
}
}  // namespace plas
