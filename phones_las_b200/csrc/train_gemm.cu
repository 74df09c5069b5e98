// Training-path dense algebra in exact fp32 (the reference trains in fp32, model_helper.py:403-417):
//   plas_gemm_f32_ex   C = alpha * op(A) op(B) + beta * C + bias, arbitrary element strides for A and B, so the one
//                      kernel serves  x @ W (forward),  dZ @ W^T (input gradients, las/ops.py:35-40 backward) and
//                      X^T @ dZ (weight gradients) directly on the TF checkpoint layout [din+U, 4U] -- no packed or
//                      transposed copies of the weights exist on the training path;  grid.z batches independent
//                      problems (per-utterance dkeys / dvalues of the attention backward).
//   plas_colsum_f32    bias gradients (column sums, fixed order).
// 128x128x16 (or 128x64x16) tiles, 256 threads, 8x8 (8x4) register micro-tile, double-buffered shared memory; operand
// tiles are fetched with 16-byte loads along whichever dimension is contiguous (scalar fallback for odd shapes).
#include "common.cuh"
#include "../../include/plas.h"

namespace plas {

constexpr int TG_BM = 128, TG_BK = 16;

struct GemmExArgs {
  long long M;
  int N, K;
  const float* A;
  long long sam, sak;
  const float* B;
  long long sbk, sbn;
  float* C;
  long long ldc;
  const float* bias;
  float alpha, beta;
  long long batch_a, batch_b, batch_c;
  int split_k;       // > 1: blockIdx.z is a K slice, partial products go to split_ws[z][M][N]
  int k_chunk;       // K elements per slice (multiple of TG_BK)
  float* split_ws;
  int a_vec, b_vec;  // operand tiles can be fetched with aligned 16-byte loads along their contiguous dimension
};

// BN = 128: 8x8 register micro-tile (64 FMA per 4 LDS.128);  BN = 64: 8x4 (narrow outputs: vocabulary, features).
template <int BN, bool VEC>
__global__ void __launch_bounds__(256, 2) gemm_f32_ex_kernel(GemmExArgs p) {
  constexpr int TN = BN / 16;  // columns per thread (8 or 4), as TN/4 groups of 4 that are 64 columns apart
  __shared__ __align__(16) float sA[2][TG_BK][TG_BM + 4];
  __shared__ __align__(16) float sB[2][TG_BK][BN + 4];
  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.y * TG_BM;
  const int n0 = blockIdx.x * BN;
  const bool split = p.split_k > 1;
  const float* __restrict__ A = p.A + (split ? 0 : (long long)blockIdx.z * p.batch_a);
  const float* __restrict__ B = p.B + (split ? 0 : (long long)blockIdx.z * p.batch_b);
  float* __restrict__ C = split ? p.split_ws + (long long)blockIdx.z * p.M * p.N : p.C + (long long)blockIdx.z * p.batch_c;
  const int k_begin = split ? blockIdx.z * p.k_chunk : 0;
  const int k_end = split ? min(p.K, k_begin + p.k_chunk) : p.K;
  const int tx = tid % 16, ty = tid / 16;  // 16 x 16 threads; thread tile = rows ty*8.., cols tx*4 (+64)
  const bool a_kfast = p.sak == 1;         // consecutive threads walk k (A row-major) or m (A^T view)
  const bool b_nfast = p.sbn == 1;
  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  float4 ra[VEC ? 2 : 1], rb[VEC ? BN / 64 : 1];
  float sa_[VEC ? 1 : 8], sb_[VEC ? 1 : BN / 16];  // scalar staging (unaligned / odd-sized operands)

  auto load_tile = [&](int k0) {
    if (VEC) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int i = tid + j * 256;
        long long gm; int gk;
        if (a_kfast) { gm = m0 + (i >> 2); gk = k0 + 4 * (i & 3); } else { gk = k0 + (i >> 5); gm = m0 + 4 * (i & 31); }
        ra[j] = (gm < p.M && gk < k_end) ? *reinterpret_cast<const float4*>(A + gm * p.sam + gk * p.sak) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = tid + j * 256;
        int r, c;
        if (a_kfast) { r = i / TG_BK; c = i % TG_BK; } else { c = i / TG_BM; r = i % TG_BM; }
        const long long gm = m0 + r;
        const int gk = k0 + c;
        sa_[j] = (gm < p.M && gk < k_end) ? A[gm * p.sam + gk * p.sak] : 0.f;
      }
    }
    if (VEC) {
#pragma unroll
      for (int j = 0; j < BN / 64; ++j) {
        const int i = tid + j * 256;
        int gn, gk;
        if (b_nfast) { gk = k0 + i / (BN / 4); gn = n0 + 4 * (i % (BN / 4)); } else { gn = n0 + (i >> 2); gk = k0 + 4 * (i & 3); }
        rb[j] = (gn < p.N && gk < k_end) ? *reinterpret_cast<const float4*>(B + (long long)gk * p.sbk + (long long)gn * p.sbn)
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
#pragma unroll
      for (int j = 0; j < BN / 16; ++j) {
        const int i = tid + j * 256;
        int r, c;  // r = n index, c = k index
        if (b_nfast) { c = i / BN; r = i % BN; } else { r = i / TG_BK; c = i % TG_BK; }
        const int gn = n0 + r;
        const int gk = k0 + c;
        sb_[j] = (gn < p.N && gk < k_end) ? B[(long long)gk * p.sbk + (long long)gn * p.sbn] : 0.f;
      }
    }
  };
  auto store_tile = [&](int buf) {
    if (VEC) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int i = tid + j * 256;
        if (a_kfast) {
          const int r = i >> 2, c = 4 * (i & 3);
          sA[buf][c][r] = ra[j].x; sA[buf][c + 1][r] = ra[j].y; sA[buf][c + 2][r] = ra[j].z; sA[buf][c + 3][r] = ra[j].w;
        } else {
          *reinterpret_cast<float4*>(&sA[buf][i >> 5][4 * (i & 31)]) = ra[j];
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = tid + j * 256;
        int r, c;
        if (a_kfast) { r = i / TG_BK; c = i % TG_BK; } else { c = i / TG_BM; r = i % TG_BM; }
        sA[buf][c][r] = sa_[j];
      }
    }
    if (VEC) {
#pragma unroll
      for (int j = 0; j < BN / 64; ++j) {
        const int i = tid + j * 256;
        if (b_nfast) {
          *reinterpret_cast<float4*>(&sB[buf][i / (BN / 4)][4 * (i % (BN / 4))]) = rb[j];
        } else {
          const int r = i >> 2, c = 4 * (i & 3);
          sB[buf][c][r] = rb[j].x; sB[buf][c + 1][r] = rb[j].y; sB[buf][c + 2][r] = rb[j].z; sB[buf][c + 3][r] = rb[j].w;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < BN / 16; ++j) {
        const int i = tid + j * 256;
        int r, c;
        if (b_nfast) { c = i / BN; r = i % BN; } else { r = i / TG_BK; c = i % TG_BK; }
        sB[buf][c][r] = sb_[j];
      }
    }
  };

  const int n_tiles = (k_end - k_begin + TG_BK - 1) / TG_BK;
  load_tile(k_begin);
  store_tile(0);
  __syncthreads();
  for (int kt = 0; kt < n_tiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < n_tiles) load_tile(k_begin + (kt + 1) * TG_BK);  // global loads in flight during the FMAs
#pragma unroll
    for (int k = 0; k < TG_BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&sA[buf][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&sA[buf][k][ty * 8 + 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[TN];
#pragma unroll
      for (int g = 0; g < TN / 4; ++g) {
        const float4 bv = *reinterpret_cast<const float4*>(&sB[buf][k][tx * 4 + 64 * g]);
        b[4 * g] = bv.x; b[4 * g + 1] = bv.y; b[4 * g + 2] = bv.z; b[4 * g + 3] = bv.w;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < n_tiles) store_tile(buf ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long gm = m0 + ty * 8 + i;
    if (gm >= p.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int gn = n0 + tx * 4 + 64 * (j / 4) + (j % 4);
      if (gn >= p.N) continue;
      if (split) {
        C[gm * p.N + gn] = acc[i][j];
        continue;
      }
      float v = p.alpha * acc[i][j];
      if (p.bias) v += p.bias[gn];
      if (p.beta != 0.f) v += p.beta * C[gm * p.ldc + gn];
      C[gm * p.ldc + gn] = v;
    }
  }
}

// C = alpha * sum_z ws[z] + beta * C + bias  (fixed summation order: deterministic split-K)
__global__ void __launch_bounds__(256) splitk_reduce_kernel(GemmExArgs p) {
  const long long total = p.M * p.N;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long m = i / p.N;
    const int n = (int)(i - m * p.N);
    float s = 0.f;
    for (int z = 0; z < p.split_k; ++z) s += p.split_ws[(long long)z * total + i];
    float v = p.alpha * s;
    if (p.bias) v += p.bias[n];
    if (p.beta != 0.f) v += p.beta * p.C[m * p.ldc + n];
    p.C[m * p.ldc + n] = v;
  }
}

// out[n] = sum_m X[m][n] (+ out[n] when accumulate), fixed summation order.  A cluster of CS_CL CTAs shares 32 columns: CTA r
// sums the rows of slice r (8 row groups per CTA), the partials meet in the shared memory of rank 0 through DSMEM and are added in
// rank order -- no scratch buffer, no atomics, 8x the CTAs of the one-CTA-per-32-columns version (78 -> ~10 us for 9504 x 1024).
constexpr int CS_CL = 8;
__global__ void __cluster_dims__(1, CS_CL, 1) __launch_bounds__(256)
    colsum_f32_kernel(const float* __restrict__ X, long long M, int N, long long ld, float* __restrict__ out, int accumulate) {
  __shared__ float part[8][33];
  __shared__ float slice[CS_CL][32];  // rank 0: the per-CTA sums, written by the peers
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + lane;
  const int rank = blockIdx.y;  // cluster spans grid.y
  const long long rows = (M + CS_CL - 1) / CS_CL;
  const long long m_lo = rank * rows, m_hi = m_lo + rows < M ? m_lo + rows : M;
  float s = 0.f;
  if (n < N)
    for (long long m = m_lo + g; m < m_hi; m += 8) s += X[m * ld + n];
  part[g][lane] = s;
  __syncthreads();
  if (g == 0) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += part[k][lane];
    uint32_t dst;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dst) : "r"(smem_u32(&slice[rank][lane])), "r"(0));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(dst), "f"(t) : "memory");
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (rank == 0 && g == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < CS_CL; ++k) t += slice[k][lane];
    out[n] = accumulate ? out[n] + t : t;
  }
}

}  // namespace plas

using namespace plas;

extern "C" int plas_gemm_f32_ex(const plas_gemm_ex_desc* d, plas_stream_t stream) {
  PLAS_REQUIRE(d && d->A && d->B && d->C, "gemm_ex: null pointer");
  PLAS_REQUIRE(d->M > 0 && d->N > 0 && d->K > 0 && d->batch >= 1, "gemm_ex: bad shape M=%lld N=%d K=%d batch=%d",
               (long long)d->M, d->N, d->K, d->batch);
  GemmExArgs a;
  a.M = d->M; a.N = d->N; a.K = d->K;
  a.A = d->A; a.sam = d->sam; a.sak = d->sak;
  a.B = d->B; a.sbk = d->sbk; a.sbn = d->sbn;
  a.C = d->C; a.ldc = d->ldc;
  a.bias = d->bias; a.alpha = d->alpha; a.beta = d->beta;
  a.batch_a = d->batch_a; a.batch_b = d->batch_b; a.batch_c = d->batch_c;
  const long long gy = (d->M + TG_BM - 1) / TG_BM;
  PLAS_REQUIRE(gy <= 65535 && d->batch <= 65535, "gemm_ex: M or batch too large");
  auto al16 = [](const void* q) { return ((uintptr_t)q & 15) == 0; };
  // 16-byte operand fetches: contiguous dimension with a multiple-of-4 extent, every row / batch start aligned
  if (d->sak == 1) a.a_vec = d->K % 4 == 0 && d->sam % 4 == 0;
  else if (d->sam == 1) a.a_vec = d->M % 4 == 0 && d->sak % 4 == 0;
  else a.a_vec = 0;
  a.a_vec = a.a_vec && al16(d->A) && d->batch_a % 4 == 0;
  if (d->sbn == 1) a.b_vec = d->N % 4 == 0 && d->sbk % 4 == 0;
  else if (d->sbk == 1) a.b_vec = d->K % 4 == 0 && d->sbn % 4 == 0;
  else a.b_vec = 0;
  a.b_vec = a.b_vec && al16(d->B) && d->batch_b % 4 == 0;
  const bool vec = a.a_vec && a.b_vec;
  const int BN = (vec && d->N > 64) ? 128 : 64;
  const long long gx = (d->N + BN - 1) / BN;
  // deterministic split-K for long reductions over few output tiles (weight gradients: K = B*T rows)
  int split = 1;
  if (d->split_ws && d->batch == 1 && d->K >= 512) {
    const long long tiles = gx * gy;
    const long long want = (2LL * num_sms() + tiles - 1) / tiles;
    const long long by_ws = (long long)(d->split_ws_bytes / ((size_t)d->M * d->N * 4));
    long long sp = want < d->K / 256 ? want : d->K / 256;
    if (sp > by_ws) sp = by_ws;
    if (sp > 64) sp = 64;
    if (sp > 1) split = (int)sp;
  }
  a.split_k = split;
  a.k_chunk = d->K;
  a.split_ws = d->split_ws;
  if (split > 1) {
    a.k_chunk = ((d->K + split - 1) / split + TG_BK - 1) / TG_BK * TG_BK;
    a.split_k = (d->K + a.k_chunk - 1) / a.k_chunk;
  }
  dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)(a.split_k > 1 ? a.split_k : d->batch));
  if (BN == 128) gemm_f32_ex_kernel<128, true><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  else if (vec) gemm_f32_ex_kernel<64, true><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  else gemm_f32_ex_kernel<64, false><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  if (a.split_k > 1) {
    const long long total = d->M * d->N;
    const int blocks = (int)((total + 255) / 256 < 4LL * num_sms() ? (total + 255) / 256 : 4LL * num_sms());
    splitk_reduce_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a);
  }
  PLAS_CUDA(cudaGetLastError());
  return PLAS_OK;
}

extern "C" int plas_colsum_f32(const float* X, int64_t M, int32_t N, int64_t ld, float* out, int32_t accumulate,
                               plas_stream_t stream) {
  PLAS_REQUIRE(X && out && M > 0 && N > 0 && ld >= N, "colsum: bad argument");
  colsum_f32_kernel<<<dim3((unsigned)((N + 31) / 32), CS_CL), 256, 0, (cudaStream_t)stream>>>(X, M, N, ld, out, accumulate);
  PLAS_CUDA(cudaGetLastError());
  return PLAS_OK;
}
