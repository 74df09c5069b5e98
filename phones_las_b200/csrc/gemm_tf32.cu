// K2 for the fp32 (training / reference-precision) path on the tensor pipe: 3xTF32 split products.
//
// The reference trains in float32 (LSTMCell matmul of las/ops.py:11-12, forward, dX and dW).  An exact-fp32 SIMT GEMM tops
// out at the FFMA pipe; tcgen05 `kind::tf32` reads fp32 words and uses their top 19 bits (10-bit mantissa).  Splitting every
// operand into hi = nearest TF32 of x and lo = nearest TF32 of x - hi gives
//     a*b = a_hi*b_hi + a_lo*b_hi + a_hi*b_lo + O(2^-21 |a||b|)
// -- fp32-level products with three tensor-core MMAs accumulated in fp32 in TMEM.  The three products are ONE GEMM over a
// concatenated contraction axis:  [A_hi | A_lo | A_hi] . [B_hi | B_hi | B_lo]^T, so the kernel below is a plain "TN" GEMM
//     C[M][N] (+)= A3[M][K3] * B3[N][K3]^T (+ bias[N])
// fed by TMA (128-byte rows = 32 fp32, SWIZZLE_128B) exactly like the bf16 kernel in gemm.cu (same byte geometry: a
// K-step of 8 tf32 = 32 bytes), and `plas_split3_f32` writes the concatenated operands (optionally transposed, which turns
// the NT / TN products of the backward pass -- dZ W^T and X^T dZ on the TF layouts -- into the same TN form).
//   warp 0: TMA producer, warp 1: MMA issuer (one elected thread), warp 2: TMEM allocator, warps 4..7: epilogue.
#include <mutex>

#include "common.cuh"
#include "tcgen05.cuh"
#include "../../include/plas.h"

namespace plas {

constexpr int T_BM = 128;
constexpr int T_BK = 32;      // fp32 elements per k block (128-byte rows)

template <int BN>
struct TfCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int A_BYTES = T_BM * T_BK * 4;
  static constexpr int B_BYTES = BN * T_BK * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  // instruction descriptor of tcgen05.mma kind::tf32: D = f32 (bit 4), A = B = TF32 (format 2 at bits 7 and 10), both
  // K-major, N >> 3 at bit 17, M >> 4 at bit 24
  static constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(T_BM >> 4) << 24);
};

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Work item = (output tile, K slice).  ksplit == 1: the epilogue writes C (+ bias, + old C).  ksplit > 1 (few output tiles, long
// contraction: the weight gradients): every slice writes its raw partial tile to scratch[slice][M][ld_s] and tf32_reduce_kernel
// adds the slices in a fixed order -- deterministic, no atomics.
template <int BN>
__global__ void __launch_bounds__(256, 1)
gemm_tf32x3_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                           const float* __restrict__ bias, float* __restrict__ C, long long M, int N, int K3, long long ldc,
                           int accumulate, int ksplit, int kper, float* __restrict__ scratch, long long ld_s) {
  using Cfg = TfCfg<BN>;
  extern __shared__ unsigned char tf_smem_raw[];
  const uint32_t raw = smem_u32(tf_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* smem = tf_smem_raw + (base - raw);
  const uint32_t bar_base = base + Cfg::STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + s); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + 8 * (2 * Cfg::STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (int)((M + T_BM - 1) / T_BM);
  const int num_n = (N + BN - 1) / BN;
  const int num_items = num_m * num_n * ksplit;
  const int num_k = (K3 + T_BK - 1) / T_BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int tile = item / ksplit, slice = item - tile * ksplit;
        const int m_blk = tile / num_n, n_blk = tile % num_n;
        const int kb1 = min(num_k, (slice + 1) * kper);
        for (int kb = slice * kper; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
          const uint32_t sa = base + stage * Cfg::STAGE_BYTES;
          tma_load_2d(sa, &tmA, kb * T_BK, m_blk * T_BM, full_bar(stage));
          tma_load_2d(sa + Cfg::A_BYTES, &tmB, kb * T_BK, n_blk * BN, full_bar(stage));
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int slice = item % ksplit;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        const int kb0 = slice * kper, kb1 = min(num_k, (slice + 1) * kper);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = base + stage * Cfg::STAGE_BYTES;
          const uint64_t adesc = umma_smem_desc(sa);
          const uint64_t bdesc = umma_smem_desc(sa + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < T_BK / 8; ++k)  // K = 8 tf32 (32 bytes) per instruction
            umma_tf32(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), Cfg::IDESC, (kb != kb0 || k != 0) ? 1u : 0u);
          umma_commit(empty_bar(stage));
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(acc));
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int ew = warp - 4;  // == warp % 4 -> TMEM lanes [32*ew, 32*ew+32)
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int tile = item / ksplit, slice = item - tile * ksplit;
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const long long row = (long long)m_blk * T_BM + ew * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN);
      const bool partial = ksplit > 1;
      float* crow = partial ? scratch + ((long long)slice * M + row) * ld_s + (long long)n_blk * BN
                            : C + row * ldc + (long long)n_blk * BN;
      const float* brow = (bias && !partial) ? bias + (long long)n_blk * BN : nullptr;
      const bool add_old = accumulate && !partial;
      const int n_left = N - n_blk * BN;  // columns of this tile that exist (N need not be a multiple of BN)
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + (uint32_t)c0, r);
        tmem_ld_wait();
        if (row < M) {
          if (c0 + 32 <= n_left) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
              if (brow) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(brow + c0 + j));
                v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
              }
              if (add_old) {
                const float4 o = *reinterpret_cast<const float4*>(crow + c0 + j);
                v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
              }
              *reinterpret_cast<float4*>(crow + c0 + j) = v;
            }
          } else {
            for (int j = 0; j < 32 && c0 + j < n_left; ++j) {
              float v = __uint_as_float(r[j]) + (brow ? __ldg(brow + c0 + j) : 0.f);
              if (add_old) v += crow[c0 + j];
              crow[c0 + j] = v;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS) : "memory");
  }
}

// C[r][c] = (accumulate ? C : 0) + bias[c] + scratch[0][r][c] + scratch[1][r][c] + ...   (fixed order)
__global__ void __launch_bounds__(256) tf32_reduce_kernel(const float* __restrict__ scratch, int ksplit, long long M, int N, long long ld_s,
                                                          const float* __restrict__ bias, float* __restrict__ C, long long ldc,
                                                          int accumulate) {
  const long long total = M * (long long)N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / N;
    const int c = (int)(i - r * N);
    float v = accumulate ? C[r * ldc + c] : 0.f;
    if (bias) v += __ldg(bias + c);
    for (int s = 0; s < ksplit; ++s) v += scratch[((long long)s * M + r) * ld_s + c];
    C[r * ldc + c] = v;
  }
}

// The split (hi = nearest TF32, lo = nearest TF32 of the exact remainder) and the operand writers.
// out row = [seg0 | seg1 | seg2], each `seg_ld` wide (zero padded past `inner`): pattern 0 = (hi, lo, hi) for the A operand,
// pattern 1 = (hi, hi, lo) for the B operand.  Plain: out[r][seg*seg_ld + c] from X[r][c] (inner = cols).  Transposed:
// out[c][seg*seg_ld + r] from X[r][c] (inner = rows), through a 32x32 shared-memory tile so both sides stay coalesced.
__device__ __forceinline__ float round_tf32(float x) {
  // nearest TF32 (10 mantissa bits), ties away from zero: kind::tf32 truncates an fp32 word to its top 19 bits, which is then exact
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = round_tf32(x);
  lo = round_tf32(x - hi);  // x - hi is exact in fp32 (|x - hi| <= 2^-11 |x|); its own rounding costs 2^-22 |x|
}

__global__ void __launch_bounds__(256) split3_kernel(const float* __restrict__ X, long long rows, int cols, long long ld,
                                                     float* __restrict__ out, long long ld_out, long long seg_ld, int pattern) {
  const long long r = blockIdx.y;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < seg_ld; c += gridDim.x * blockDim.x) {
    float hi = 0.f, lo = 0.f;
    if (c < cols) split_tf32(X[r * ld + c], hi, lo);
    float* o = out + r * ld_out + c;
    o[0] = hi;
    o[seg_ld] = pattern == 0 ? lo : hi;
    o[2 * seg_ld] = pattern == 0 ? hi : lo;
  }
}

__global__ void __launch_bounds__(256) split3_t_kernel(const float* __restrict__ X, long long rows, int cols, long long ld,
                                                       float* __restrict__ out, long long ld_out, long long seg_ld, int pattern) {
  __shared__ float tile[32][33];
  const long long r0 = (long long)blockIdx.y * 32;
  const int c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const long long r = r0 + i;
    const int c = c0 + tx;
    tile[i][tx] = (r < rows && c < cols) ? X[r * ld + c] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i;          // output row
    const long long r = r0 + tx;   // output column inside a segment
    if (c < cols && r < seg_ld) {
      float hi, lo;
      split_tf32(tile[tx][i], hi, lo);
      float* o = out + (long long)c * ld_out + r;
      o[0] = hi;
      o[seg_ld] = pattern == 0 ? lo : hi;
      o[2 * seg_ld] = pattern == 0 ? hi : lo;
    }
  }
}

inline int make_map_f32(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_err(PLAS_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)T_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(PLAS_ECUDA, "cuTensorMapEncodeTiled(f32) failed with CUresult %d", (int)r);
  return PLAS_OK;
}

}  // namespace plas

using namespace plas;

extern "C" int plas_split3_f32(const float* X, int64_t rows, int32_t cols, int64_t ld, float* out, int64_t ld_out, int64_t seg_ld,
                               int32_t pattern, int32_t transpose, plas_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PLAS_REQUIRE(X && out && rows > 0 && cols > 0 && ld >= cols, "split3: bad arguments");
  PLAS_REQUIRE(pattern == 0 || pattern == 1, "split3: pattern %d", pattern);
  PLAS_REQUIRE(seg_ld % 4 == 0 && ld_out >= 3 * seg_ld, "split3: seg_ld=%lld must be a multiple of 4 and ld_out >= 3*seg_ld", (long long)seg_ld);
  if (!transpose) {
    PLAS_REQUIRE(seg_ld >= cols, "split3: seg_ld < cols");
    PLAS_REQUIRE(rows <= 2147483647LL, "split3: too many rows");
    long long done = 0;
    while (done < rows) {  // grid.y <= 65535
      const long long chunk = rows - done < 65535 ? rows - done : 65535;
      dim3 grid((unsigned)((seg_ld + 255) / 256 < 8 ? (seg_ld + 255) / 256 : 8), (unsigned)chunk);
      split3_kernel<<<grid, 256, 0, stream>>>(X + done * ld, chunk, cols, ld, out + done * ld_out, ld_out, seg_ld, pattern);
      done += chunk;
    }
  } else {
    PLAS_REQUIRE(seg_ld >= rows, "split3(transposed): seg_ld < rows");
    const long long gy = (seg_ld + 31) / 32;
    PLAS_REQUIRE(gy <= 65535, "split3(transposed): too many rows");
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)gy);
    split3_t_kernel<<<grid, 256, 0, stream>>>(X, rows, cols, ld, out, ld_out, seg_ld, pattern);
  }
  PLAS_CUDA(cudaGetLastError());
  return PLAS_OK;
}

// K slices for a problem: 1 when the output tiles alone fill the GPU, otherwise enough slices to do so, each at least 16 k blocks
static void tf32_plan(long long M, int N, int K3, int* bn, int* ksplit, int* kper) {
  const int sms = num_sms() > 0 ? num_sms() : 148;
  const int num_k = (K3 + T_BK - 1) / T_BK;
  const long long m_tiles = (M + T_BM - 1) / T_BM;
  // 256-wide tiles halve the MMA count per flop (an SS-mode MMA costs ~140 clk whatever its N): use them when they still fill the GPU
  *bn = (N % 256 == 0 && m_tiles * (N / 256) >= sms) ? 256 : 128;
  const long long tiles = m_tiles * ((N + *bn - 1) / *bn);
  int ks = 1;
  if (tiles * 2 <= sms) {
    ks = (int)((sms + tiles - 1) / tiles);
    if (ks > num_k / 16) ks = num_k / 16;
    if (ks > 32) ks = 32;
    if (ks < 1) ks = 1;
  }
  *kper = (num_k + ks - 1) / ks;
  *ksplit = (num_k + *kper - 1) / *kper;  // every slice owns at least one k block
}

extern "C" size_t plas_gemm_tf32x3_scratch_bytes(int64_t M, int32_t N, int32_t K3) {
  int bn, ks, kper;
  tf32_plan(M, N, K3, &bn, &ks, &kper);
  return ks > 1 ? (size_t)ks * (size_t)M * (size_t)((N + 3) & ~3) * 4 : 0;
}

template <int BN>
static int launch_tf32(const CUtensorMap& tmA, const CUtensorMap& tmB, const float* bias, float* C, long long M, int N, int K3, long long ldc,
                       int accumulate, int ksplit, int kper, float* scratch, long long ld_s, cudaStream_t stream) {
  using Cfg = TfCfg<BN>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(gemm_tf32x3_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
  });
  PLAS_CUDA(attr_err);
  const long long items = ((M + T_BM - 1) / T_BM) * ((N + BN - 1) / BN) * ksplit;
  const int sms = num_sms() > 0 ? num_sms() : 148;
  const int grid = (int)(items < sms ? items : sms);
  gemm_tf32x3_tcgen05_kernel<BN><<<grid, 256, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, bias, C, M, N, K3, ldc, accumulate, ksplit, kper, scratch, ld_s);
  PLAS_CUDA(cudaGetLastError());
  return PLAS_OK;
}

extern "C" int plas_gemm_tf32x3_tn(const float* A3, int64_t M, int32_t K3, int64_t lda, const float* B3, int32_t N, int64_t ldb,
                                   const float* bias, float* C, int64_t ldc, int32_t accumulate, void* scratch, size_t scratch_bytes,
                                   plas_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PLAS_REQUIRE(A3 && B3 && C, "gemm_tf32x3: null pointer");
  PLAS_REQUIRE(M > 0 && K3 > 0 && N > 0, "gemm_tf32x3: M=%lld K3=%d N=%d", (long long)M, K3, N);
  PLAS_REQUIRE(lda % 4 == 0 && ldb % 4 == 0 && ldc % 4 == 0 && lda >= K3 && ldb >= K3 && ldc >= N, "gemm_tf32x3: leading dimensions");
  PLAS_REQUIRE(((uintptr_t)A3 % 16) == 0 && ((uintptr_t)B3 % 16) == 0 && ((uintptr_t)C % 16) == 0 && (!bias || ((uintptr_t)bias % 16) == 0),
               "gemm_tf32x3: pointers must be 16-byte aligned");
  int bn, ksplit, kper;
  tf32_plan(M, N, K3, &bn, &ksplit, &kper);
  const long long ld_s = (N + 3) & ~3;
  if (ksplit > 1) {
    const size_t need = (size_t)ksplit * (size_t)M * (size_t)ld_s * 4;
    PLAS_REQUIRE(scratch && scratch_bytes >= need && ((uintptr_t)scratch % 16) == 0, "gemm_tf32x3: split-K scratch %zu < %zu", scratch_bytes, need);
  }
  CUtensorMap tmA, tmB;
  int rc = make_map_f32(&tmA, A3, M, K3, lda, T_BM);
  if (rc) return rc;
  rc = make_map_f32(&tmB, B3, N, K3, ldb, bn);
  if (rc) return rc;
  rc = bn == 256 ? launch_tf32<256>(tmA, tmB, bias, C, M, N, K3, ldc, accumulate, ksplit, kper, (float*)scratch, ld_s, stream)
                 : launch_tf32<128>(tmA, tmB, bias, C, M, N, K3, ldc, accumulate, ksplit, kper, (float*)scratch, ld_s, stream);
  if (rc) return rc;
  if (ksplit > 1) {
    const long long total = M * (long long)N;
    const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
    tf32_reduce_kernel<<<blocks, 256, 0, stream>>>((const float*)scratch, ksplit, M, N, ld_s, bias, C, ldc, accumulate);
    PLAS_CUDA(cudaGetLastError());
  }
  return PLAS_OK;
}
