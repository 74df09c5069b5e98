// K2 -- time-parallel projections  C[M][N] = A[M][K] * Wt[N][K]^T + bias[N].
//
// bf16 path (plas_gemm_bf16): persistent warp-specialised tcgen05 kernel for sm_100a.
//   warp 0      : TMA producer   (cp.async.bulk.tensor 2D, 128B swizzle, mbarrier complete_tx)
//   warp 1      : MMA issuer     (one thread issues tcgen05.mma cta_group::1 kind::f16,
//                                 M=128, N=BN, K=16 per instruction; accumulators in TMEM)
//   warp 2      : TMEM allocator (2 accumulator stages x BN columns)
//   warps 4..7  : epilogue       (tcgen05.ld 32x32b -> +bias -> bf16 -> 16-byte global stores)
//   warp 3      : work scheduler (cluster launch control: clusterlaunchcontrol.try_cancel)
// Three pipelines: smem full/empty (TMA<->MMA), TMEM full/empty (MMA<->epilogue), and the work ring (scheduler -> the
// other roles).  The grid has ONE CTA PER WORK UNIT (up to four n-blocks of one m-block); a running CTA takes over units
// whose CTAs have not started yet by cancelling their launch, so the kernel is persistent on however many SMs it gets --
// all 148 when it runs alone, the 52-84 a co-resident recurrence of the other stream's batch leaves otherwise (a static
// round-robin grid cannot finish before its last CTA has found an SM).  K and M tails rely on TMA out-of-bounds zero
// fill; N must be a multiple of 128.
//
// f32 path (plas_gemm_f32): exact-fp32 SIMT kernel, used only by the reference-precision mode.
#include <mutex>

#include "common.cuh"
#include "tcgen05.cuh"
#include "../../include/plas.h"

namespace plas {

// ---------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------
constexpr int G_BM = 128;
constexpr int G_BK = 64;
constexpr int G_SCHED = 4;  // depth of the work ring: units a CTA may have claimed ahead of the one it is working on

// ---- cluster launch control: take over the work of a CTA of this grid that has not been launched yet -------------------
__device__ __forceinline__ void clc_try_cancel(uint32_t resp, uint32_t bar) {
  asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.b128 [%0], [%1];" ::"r"(resp), "r"(bar)
               : "memory");
}
// -> true and the cancelled CTA's blockIdx.x when a launch was cancelled, false when every CTA of the grid has started
__device__ __forceinline__ bool clc_read(uint32_t resp, int& bx) {
  uint32_t valid, x = 0, y, z;
  asm volatile(
      "{\n\t.reg .pred p1;\n\t.reg .b128 r;\n\t"
      "ld.shared.b128 r, [%4];\n\t"
      "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p1, r;\n\t"
      "selp.u32 %3, 1, 0, p1;\n\t"
      "@p1 clusterlaunchcontrol.query_cancel.get_first_ctaid.v4.b32.b128 {%0, %1, %2, _}, r;\n\t}"
      : "=r"(x), "=r"(y), "=r"(z), "=r"(valid)
      : "r"(resp)
      : "memory");
  (void)y;
  (void)z;
  bx = (int)x;
  return valid != 0;
}

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int A_BYTES = G_BM * G_BK * 2;
  static constexpr int B_BYTES = BN * G_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 512 /*barriers, work ring*/;
  // instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 at bit 17, M>>4 at bit 24
  static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                    ((uint32_t)(G_BM >> 4) << 24);
};

template <int BN, bool OUT_F32>
__global__ void __launch_bounds__(256, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const float* __restrict__ bias, void* __restrict__ Cv, long long M, int N,
                         int K, long long ldc, int tpu) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ unsigned char gemm_smem_raw[];
  const uint32_t raw = smem_u32(gemm_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* smem = gemm_smem_raw + (base - raw);
  const uint32_t bar_base = base + Cfg::STAGES * Cfg::STAGE_BYTES;
  // barriers: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2]; then the TMEM base slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + s); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + 8 * (2 * Cfg::STAGES + 4));
  // work ring: slot i = response of the i-th try_cancel (16 bytes) + "response landed" / "every role has read it" barriers
  auto wfull_bar = [&](int s) { return bar_base + 144u + 8u * s; };
  auto wempty_bar = [&](int s) { return bar_base + 144u + 8u * (G_SCHED + s); };
  auto wresp = [&](int s) { return bar_base + 272u + 16u * s; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (int)((M + G_BM - 1) / G_BM);
  const int num_n = N / BN;
  const int num_k = (K + G_BK - 1) / G_BK;
  const int upm = num_n / tpu;  // work units per m-block; unit u = n-blocks (u % upm) * tpu ... + tpu - 1 of m-block u / upm
  (void)num_m;
  // every role walks the same unit sequence: its own blockIdx.x first, then whatever the scheduler's cancellations return
  auto next_unit = [&](int it, int& unit, bool leader, bool whole_warp) -> bool {
    const int slot = it % G_SCHED;
    mbar_wait(wfull_bar(slot), (uint32_t)((it / G_SCHED) & 1));
    const bool ok = clc_read(wresp(slot), unit);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // this read vs the next asynchronous write of the slot
    if (whole_warp) __syncwarp();  // every lane has read the response before the leader releases the slot
    if (leader) mbar_arrive(wempty_bar(slot));
    return ok;
  };

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
    for (int s = 0; s < G_SCHED; ++s) {
      mbar_init(wfull_bar(s), 1);
      mbar_init(wempty_bar(s), 6);  // TMA producer, MMA issuer, four epilogue warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(Cfg::TMEM_COLS)
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
      int unit = blockIdx.x, it = 0;
      do {
        const int m_blk = unit / upm, nb0 = (unit % upm) * tpu;
        for (int sub = 0; sub < tpu; ++sub) {
          for (int kb = 0; kb < num_k; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            mbar_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
            const uint32_t sa = base + stage * Cfg::STAGE_BYTES;
            tma_load_2d(sa, &tmA, kb * G_BK, m_blk * G_BM, full_bar(stage));
            tma_load_2d(sa + Cfg::A_BYTES, &tmB, kb * G_BK, (nb0 + sub) * BN, full_bar(stage));
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      } while (next_unit(it++, unit, true, false));
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int unit = blockIdx.x, it = 0;
      do {
        for (int sub = 0; sub < tpu; ++sub) {
          mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
          for (int kb = 0; kb < num_k; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t sa = base + stage * Cfg::STAGE_BYTES;
            const uint64_t adesc = umma_smem_desc(sa);
            const uint64_t bdesc = umma_smem_desc(sa + Cfg::A_BYTES);
#pragma unroll
            for (int k = 0; k < G_BK / 16; ++k)
              umma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), Cfg::IDESC,
                        (kb | k) != 0 ? 1u : 0u);
            umma_commit(empty_bar(stage));
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
          }
          umma_commit(tfull_bar(acc));
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1u;
        }
      } while (next_unit(it++, unit, true, false));
    }
    __syncwarp();
  } else if (warp == 3) {
    // ===== work scheduler: one cancellation request at a time (a request after a failed one is undefined); it runs up to
    // G_SCHED units ahead of the roles that consume the responses =====
    if (elect_one()) {
      int it = 0, dummy;
      while (true) {
        const int slot = it % G_SCHED;
        if (it >= G_SCHED) mbar_wait(wempty_bar(slot), (uint32_t)(((it / G_SCHED) - 1) & 1));
        mbar_expect_tx(wfull_bar(slot), 16);
        clc_try_cancel(wresp(slot), wfull_bar(slot));
        mbar_wait(wfull_bar(slot), (uint32_t)((it / G_SCHED) & 1));
        const bool ok = clc_read(wresp(slot), dummy);
        ++it;
        if (!ok) break;
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int ew = warp - 4;  // == warp % 4 -> TMEM lanes [32*ew, 32*ew+32)
    int acc = 0;
    uint32_t acc_phase = 0;
    int unit = blockIdx.x, it = 0;
    do {
     const int m_blk = unit / upm, nb0 = (unit % upm) * tpu;
     for (int sub = 0; sub < tpu; ++sub) {
      const int n_blk = nb0 + sub;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const long long row = (long long)m_blk * G_BM + ew * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN);
      __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(Cv) + row * ldc + (long long)n_blk * BN;
      float* crow32 = reinterpret_cast<float*>(Cv) + row * ldc + (long long)n_blk * BN;
      const float* brow = bias ? bias + (long long)n_blk * BN : nullptr;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + (uint32_t)c0, r);
        tmem_ld_wait();
        if (row < M && OUT_F32) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 v;
            v.x = __uint_as_float(r[j]) + (brow ? __ldg(brow + c0 + j) : 0.f);
            v.y = __uint_as_float(r[j + 1]) + (brow ? __ldg(brow + c0 + j + 1) : 0.f);
            v.z = __uint_as_float(r[j + 2]) + (brow ? __ldg(brow + c0 + j + 2) : 0.f);
            v.w = __uint_as_float(r[j + 3]) + (brow ? __ldg(brow + c0 + j + 3) : 0.f);
            *reinterpret_cast<float4*>(crow32 + c0 + j) = v;
          }
        } else if (row < M) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint32_t pk[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float v0 = __uint_as_float(r[j + 2 * q]);
              float v1 = __uint_as_float(r[j + 2 * q + 1]);
              if (brow) {
                v0 += __ldg(brow + c0 + j + 2 * q);
                v1 += __ldg(brow + c0 + j + 2 * q + 1);
              }
              __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
              pk[q] = *reinterpret_cast<uint32_t*>(&h);
            }
            *reinterpret_cast<uint4*>(crow + c0 + j) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
     }
     __syncwarp();
    } while (next_unit(it++, unit, lane == 0, true));
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------
// exact fp32 SIMT GEMM (reference-precision mode)
// ---------------------------------------------------------------------------------------
constexpr int S_BM = 64, S_BN = 64, S_BK = 16;

__global__ void __launch_bounds__(256) gemm_f32_kernel(const float* __restrict__ A, long long M, int K,
                                                       long long lda, const float* __restrict__ Wt, int N,
                                                       long long ldw, const float* __restrict__ bias,
                                                       float* __restrict__ C, long long ldc) {
  __shared__ float sA[S_BK][S_BM + 4];
  __shared__ float sB[S_BK][S_BN + 4];
  const long long m0 = (long long)blockIdx.y * S_BM;
  const int n0 = blockIdx.x * S_BN;
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;  // 16x16 threads, 4x4 outputs each
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += S_BK) {
    for (int i = tid; i < S_BM * S_BK; i += 256) {
      const int r = i / S_BK, c = i % S_BK;
      const long long gm = m0 + r;
      const int gk = k0 + c;
      sA[c][r] = (gm < M && gk < K) ? A[gm * lda + gk] : 0.f;
      const int gn = n0 + r;
      sB[c][r] = (gn < N && gk < K) ? Wt[(long long)gn * ldw + gk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < S_BK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn < N) C[gm * ldc + gn] = acc[i][j] + (bias ? bias[gn] : 0.f);
    }
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
template <int BN, bool OUT_F32>
static int launch_gemm_bf16(const void* A, long long M, int K, long long lda, const void* Wt, int N, long long ldw,
                            const float* bias, void* C, long long ldc, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  CUtensorMap tmA, tmB;
  int rc = make_map_bf16(&tmA, A, M, K, lda, G_BM);
  if (rc) return rc;
  rc = make_map_bf16(&tmB, Wt, N, K, ldw, BN);
  if (rc) return rc;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN, OUT_F32>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
  });
  PLAS_CUDA(attr_err);
  // one CTA per work unit = up to 4 consecutive n-blocks of one m-block (a unit must outlast a cancellation round trip);
  // CTAs that find an SM take over the units of those that have not started (cluster launch control)
  const int num_n = N / BN;
  const int tpu = num_n % 4 == 0 ? 4 : (num_n % 2 == 0 ? 2 : 1);
  const long long units = ((M + G_BM - 1) / G_BM) * (num_n / tpu);
  PLAS_REQUIRE(units <= 0x7fffffffLL, "gemm_bf16: %lld work units exceed the grid limit", units);
  gemm_bf16_tcgen05_kernel<BN, OUT_F32><<<(unsigned)units, 256, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, bias, C, M, N, K, ldc, tpu);
  PLAS_CUDA(cudaGetLastError());
  return PLAS_OK;
}

}  // namespace plas

using namespace plas;

static int gemm_bf16_dispatch(const void* A, int64_t M, int32_t K, int64_t lda, const void* Wt, int32_t N,
                              int64_t ldw, const float* bias, void* C, int64_t ldc, bool out_f32, plas_stream_t stream) {
  PLAS_REQUIRE(A && Wt && C, "gemm_bf16: null pointer");
  PLAS_REQUIRE(M > 0 && K > 0 && N > 0, "gemm_bf16: M=%lld K=%d N=%d", (long long)M, K, N);
  PLAS_REQUIRE(N % 128 == 0, "gemm_bf16: N=%d must be a multiple of 128 (pad the packed weights)", N);
  PLAS_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && ldc % 8 == 0, "gemm_bf16: lda/ldw/ldc must be multiples of 8");
  (void)out_f32;
  PLAS_REQUIRE(lda >= K && ldw >= K && ldc >= N, "gemm_bf16: leading dimension too small");
  PLAS_REQUIRE(((uintptr_t)A % 16) == 0 && ((uintptr_t)Wt % 16) == 0 && ((uintptr_t)C % 16) == 0,
               "gemm_bf16: pointers must be 16-byte aligned");
  if (out_f32) {
    if (N % 256 == 0)
      return launch_gemm_bf16<256, true>(A, M, K, lda, Wt, N, ldw, bias, C, ldc, (cudaStream_t)stream);
    return launch_gemm_bf16<128, true>(A, M, K, lda, Wt, N, ldw, bias, C, ldc, (cudaStream_t)stream);
  }
  if (N % 256 == 0)
    return launch_gemm_bf16<256, false>(A, M, K, lda, Wt, N, ldw, bias, C, ldc, (cudaStream_t)stream);
  return launch_gemm_bf16<128, false>(A, M, K, lda, Wt, N, ldw, bias, C, ldc, (cudaStream_t)stream);
}

extern "C" int plas_gemm_bf16(const void* A, int64_t M, int32_t K, int64_t lda, const void* Wt, int32_t N,
                              int64_t ldw, const float* bias, void* C, int64_t ldc, plas_stream_t stream) {
  return gemm_bf16_dispatch(A, M, K, lda, Wt, N, ldw, bias, C, ldc, false, stream);
}

extern "C" int plas_gemm_bf16_f32out(const void* A, int64_t M, int32_t K, int64_t lda, const void* Wt, int32_t N,
                                     int64_t ldw, const float* bias, float* C, int64_t ldc, plas_stream_t stream) {
  return gemm_bf16_dispatch(A, M, K, lda, Wt, N, ldw, bias, C, ldc, true, stream);
}

extern "C" int plas_gemm_f32(const float* A, int64_t M, int32_t K, int64_t lda, const float* Wt, int32_t N,
                             int64_t ldw, const float* bias, float* C, int64_t ldc, plas_stream_t stream) {
  PLAS_REQUIRE(A && Wt && C, "gemm_f32: null pointer");
  PLAS_REQUIRE(M > 0 && K > 0 && N > 0 && lda >= K && ldw >= K && ldc >= N, "gemm_f32: bad shape");
  const long long gy = (M + S_BM - 1) / S_BM;
  PLAS_REQUIRE(gy <= 65535LL * 32768, "gemm_f32: M too large");
  // grid.y is limited to 65535: fold big M into multiple launches
  long long m_done = 0;
  while (m_done < M) {
    long long m_chunk = M - m_done;
    if (m_chunk > 65535LL * S_BM) m_chunk = 65535LL * S_BM;
    dim3 grid((N + S_BN - 1) / S_BN, (unsigned)((m_chunk + S_BM - 1) / S_BM));
    gemm_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A + m_done * lda, m_chunk, K, lda, Wt, N, ldw, bias,
                                                            C + m_done * ldc, ldc);
    PLAS_CUDA(cudaGetLastError());
    m_done += m_chunk;
  }
  return PLAS_OK;
}
