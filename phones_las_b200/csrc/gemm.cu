// K2 -- time-parallel projections  C[M][N] = A[M][K] * Wt[N][K]^T + bias[N].
//
// bf16 path (plas_gemm_bf16): persistent warp-specialised tcgen05 kernel for sm_100a.
//   warp 0      : TMA producer   (cp.async.bulk.tensor 2D, 128B swizzle, mbarrier complete_tx)
//   warp 1      : MMA issuer     (one thread issues tcgen05.mma cta_group::1 kind::f16,
//                                 M=128, N=BN, K=16 per instruction; accumulators in TMEM)
//   warp 2      : TMEM allocator (2 accumulator stages x BN columns)
//   warps 4..7  : epilogue       (tcgen05.ld 32x32b -> +bias -> bf16 -> 16-byte global stores)
// Three pipelines: smem full/empty (TMA<->MMA), TMEM full/empty (MMA<->epilogue), and a static
// round-robin tile schedule over a grid of min(#tiles, #SMs) CTAs.  K and M tails rely on TMA
// out-of-bounds zero fill; N must be a multiple of 128.
//
// f32 path (plas_gemm_f32): exact-fp32 SIMT kernel, used only by the reference-precision mode.
#include <cuda.h>

#include <mutex>

#include "common.cuh"
#include "../../include/plas.h"

namespace plas {

// ---------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled shared-memory matrix descriptor (rows of 64 bf16 = 128 B; 8-row
// swizzle atoms 1024 B apart).  Matches what TMA writes with CU_TENSOR_MAP_SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);  // start address
  d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                    // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}

constexpr int G_BM = 128;
constexpr int G_BK = 64;

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int A_BYTES = G_BM * G_BK * 2;
  static constexpr int B_BYTES = BN * G_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  // instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 at bit 17, M>>4 at bit 24
  static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                    ((uint32_t)(G_BM >> 4) << 24);
};

template <int BN>
__global__ void __launch_bounds__(256, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const float* __restrict__ bias, __nv_bfloat16* __restrict__ C, long long M, int N,
                         int K, long long ldc) {
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

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (int)((M + G_BM - 1) / G_BM);
  const int num_n = N / BN;
  const int num_tiles = num_m * num_n;
  const int num_k = (K + G_BK - 1) / G_BK;

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
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / num_n, n_blk = tile % num_n;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
          const uint32_t sa = base + stage * Cfg::STAGE_BYTES;
          tma_load_2d(sa, &tmA, kb * G_BK, m_blk * G_BM, full_bar(stage));
          tma_load_2d(sa + Cfg::A_BYTES, &tmB, kb * G_BK, n_blk * BN, full_bar(stage));
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
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
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int ew = warp - 4;  // == warp % 4 -> TMEM lanes [32*ew, 32*ew+32)
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const long long row = (long long)m_blk * G_BM + ew * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN);
      __nv_bfloat16* crow = C + row * ldc + (long long)n_blk * BN;
      const float* brow = bias ? bias + (long long)n_blk * BN : nullptr;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + (uint32_t)c0, r);
        tmem_ld_wait();
        if (row < M) {
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
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

static int make_map_bf16(CUtensorMap* map, const void* ptr, long long rows, int cols, long long ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_err(PLAS_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)G_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(PLAS_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return PLAS_OK;
}

template <int BN>
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
    attr_err = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Cfg::SMEM_BYTES);
  });
  PLAS_CUDA(attr_err);
  const long long tiles = ((M + G_BM - 1) / G_BM) * (N / BN);
  const int sms = num_sms() > 0 ? num_sms() : 148;
  const int grid = (int)(tiles < sms ? tiles : sms);
  gemm_bf16_tcgen05_kernel<BN><<<grid, 256, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, bias, (__nv_bfloat16*)C, M, N, K, ldc);
  PLAS_CUDA(cudaGetLastError());
  return PLAS_OK;
}

}  // namespace plas

using namespace plas;

extern "C" int plas_gemm_bf16(const void* A, int64_t M, int32_t K, int64_t lda, const void* Wt, int32_t N,
                              int64_t ldw, const float* bias, void* C, int64_t ldc, plas_stream_t stream) {
  PLAS_REQUIRE(A && Wt && C, "gemm_bf16: null pointer");
  PLAS_REQUIRE(M > 0 && K > 0 && N > 0, "gemm_bf16: M=%lld K=%d N=%d", (long long)M, K, N);
  PLAS_REQUIRE(N % 128 == 0, "gemm_bf16: N=%d must be a multiple of 128 (pad the packed weights)", N);
  PLAS_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && ldc % 8 == 0, "gemm_bf16: lda/ldw/ldc must be multiples of 8");
  PLAS_REQUIRE(lda >= K && ldw >= K && ldc >= N, "gemm_bf16: leading dimension too small");
  PLAS_REQUIRE(((uintptr_t)A % 16) == 0 && ((uintptr_t)Wt % 16) == 0 && ((uintptr_t)C % 16) == 0,
               "gemm_bf16: pointers must be 16-byte aligned");
  if (N % 256 == 0)
    return launch_gemm_bf16<256>(A, M, K, lda, Wt, N, ldw, bias, C, ldc, (cudaStream_t)stream);
  return launch_gemm_bf16<128>(A, M, K, lda, Wt, N, ldw, bias, C, ldc, (cudaStream_t)stream);
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
