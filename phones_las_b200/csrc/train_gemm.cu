// Training-path dense algebra in exact fp32 (the reference trains in fp32, model_helper.py:403-417):
//   plas_gemm_f32_ex   C = alpha * op(A) op(B) + beta * C + bias, arbitrary element strides for A and B, so the one
//                      kernel serves  x @ W (forward),  dZ @ W^T (input gradients, las/ops.py:35-40 backward) and
//                      X^T @ dZ (weight gradients) directly on the TF checkpoint layout [din+U, 4U] -- no packed or
//                      transposed copies of the weights exist on the training path;  grid.z batches independent
//                      problems (per-utterance dkeys / dvalues of the attention backward).
//   plas_colsum_f32    bias gradients (column sums, fixed order).
// 128x64x16 tiles, 256 threads, 8x4 register micro-tile, tile loads walk whichever operand dimension is contiguous.
#include "common.cuh"
#include "../../include/plas.h"

namespace plas {

constexpr int TG_BM = 128, TG_BN = 64, TG_BK = 16;

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
};

__global__ void __launch_bounds__(256) gemm_f32_ex_kernel(GemmExArgs p) {
  __shared__ __align__(16) float sA[2][TG_BK][TG_BM + 4];
  __shared__ __align__(16) float sB[2][TG_BK][TG_BN + 4];
  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.y * TG_BM;
  const int n0 = blockIdx.x * TG_BN;
  const bool split = p.split_k > 1;
  const float* __restrict__ A = p.A + (split ? 0 : (long long)blockIdx.z * p.batch_a);
  const float* __restrict__ B = p.B + (split ? 0 : (long long)blockIdx.z * p.batch_b);
  float* __restrict__ C = split ? p.split_ws + (long long)blockIdx.z * p.M * p.N : p.C + (long long)blockIdx.z * p.batch_c;
  const int k_begin = split ? blockIdx.z * p.k_chunk : 0;
  const int k_end = split ? min(p.K, k_begin + p.k_chunk) : p.K;
  const int tx = tid % 16, ty = tid / 16;  // 16 x 16 threads; thread tile = rows ty*8.., cols tx*4..
  const bool a_kfast = p.sak == 1;         // consecutive threads walk k (A row-major) or m (A^T view)
  const bool b_nfast = p.sbn == 1;
  float acc[8][4] = {};
  float ra[8], rb[4];

  auto load_tile = [&](int k0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = tid + j * 256;
      int r, c;
      if (a_kfast) { r = i / TG_BK; c = i % TG_BK; } else { c = i / TG_BM; r = i % TG_BM; }
      const long long gm = m0 + r;
      const int gk = k0 + c;
      ra[j] = (gm < p.M && gk < k_end) ? A[gm * p.sam + gk * p.sak] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = tid + j * 256;
      int r, c;  // r = n index, c = k index
      if (b_nfast) { c = i / TG_BN; r = i % TG_BN; } else { r = i / TG_BK; c = i % TG_BK; }
      const int gn = n0 + r;
      const int gk = k0 + c;
      rb[j] = (gn < p.N && gk < k_end) ? B[gk * p.sbk + gn * p.sbn] : 0.f;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = tid + j * 256;
      int r, c;
      if (a_kfast) { r = i / TG_BK; c = i % TG_BK; } else { c = i / TG_BM; r = i % TG_BM; }
      sA[buf][c][r] = ra[j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = tid + j * 256;
      int r, c;
      if (b_nfast) { c = i / TG_BN; r = i % TG_BN; } else { r = i / TG_BK; c = i % TG_BK; }
      sB[buf][c][r] = rb[j];
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
      const float4 b0 = *reinterpret_cast<const float4*>(&sB[buf][k][tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < n_tiles) store_tile(buf ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long gm = m0 + ty * 8 + i;
    if (gm >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
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

// out[n] = sum_m X[m][n] (+ out[n] when accumulate): one CTA per 32 columns, 8 row groups, fixed order
__global__ void __launch_bounds__(256) colsum_f32_kernel(const float* __restrict__ X, long long M, int N, long long ld,
                                                         float* __restrict__ out, int accumulate) {
  __shared__ float part[8][33];
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (n < N)
    for (long long m = g; m < M; m += 8) s += X[m * ld + n];
  part[g][lane] = s;
  __syncthreads();
  if (g == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += part[k][lane];
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
  const long long gx = (d->N + TG_BN - 1) / TG_BN;
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
  gemm_f32_ex_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
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
  colsum_f32_kernel<<<(N + 31) / 32, 256, 0, (cudaStream_t)stream>>>(X, M, N, ld, out, accumulate);
  PLAS_CUDA(cudaGetLastError());
  return PLAS_OK;
}
