// K5 -- loss forward passes of the model function (reference model_helper.py:20-130, 347-358):
//   sequence softmax cross-entropy  (tf.contrib.seq2seq.sequence_loss, model_helper.py:30,75)
//   sigmoid cross-entropy over binary features (model_helper.py:81-95)
//   CTC negative log-likelihood, blank = 0  (tf.nn.ctc_loss_v2 dense-label path, model_helper.py:355-356)
// All sums are reduced in a fixed order (per-token values to HBM, then one block) so results are deterministic.
#include "common.cuh"
#include "../../include/plas.h"

namespace plas {

// warp per token: ce[b,t] = logsumexp(logits[b,t,:]) - logits[b,t,target]
__global__ void seq_ce_token_kernel(const float* __restrict__ logits, const int* __restrict__ targets, long long n_tok,
                                    int V, float* __restrict__ ce) {
  const int lane = threadIdx.x & 31;
  const long long tok = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tok >= n_tok) return;
  const float* row = logits + tok * V;
  float m = -INFINITY;
  for (int v = lane; v < V; v += 32) m = fmaxf(m, row[v]);
  m = warp_max(m);
  float s = 0.f;
  for (int v = lane; v < V; v += 32) s += expf(row[v] - m);
  s = warp_sum(s);
  if (lane == 0) {
    const int tg = max(0, min(targets[tok], V - 1));
    ce[tok] = (m + logf(s)) - row[tg];
  }
}

// warp per token: mean over n features of max(x,0) - x*z + log1p(exp(-|x|))
__global__ void sigmoid_ce_token_kernel(const float* __restrict__ logits, const float* __restrict__ labels,
                                        long long n_tok, int n, float* __restrict__ ce) {
  const int lane = threadIdx.x & 31;
  const long long tok = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tok >= n_tok) return;
  const float* x = logits + tok * n;
  const float* z = labels + tok * n;
  float s = 0.f;
  for (int k = lane; k < n; k += 32) s += fmaxf(x[k], 0.f) - x[k] * z[k] + log1pf(expf(-fabsf(x[k])));
  s = warp_sum(s);
  if (lane == 0) ce[tok] = s / (float)n;
}

// out[0] = sum(ce*w) / (sum(w) + 1e-12), out[1] = sum(ce*w), out[2] = sum(w); one block, fixed order, f64 accumulate
__global__ void weighted_mean_kernel(const float* __restrict__ ce, const float* __restrict__ w, long long n_tok,
                                     float* __restrict__ out) {
  __shared__ double s_num[256], s_den[256];
  double num = 0.0, den = 0.0;
  for (long long i = threadIdx.x; i < n_tok; i += 256) {
    const double wi = w ? (double)w[i] : 1.0;
    num += (double)ce[i] * wi;
    den += wi;
  }
  s_num[threadIdx.x] = num;
  s_den[threadIdx.x] = den;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      s_num[threadIdx.x] += s_num[threadIdx.x + o];
      s_den[threadIdx.x] += s_den[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[0] = (float)(s_num[0] / (s_den[0] + 1e-12));
    out[1] = (float)s_num[0];
    out[2] = (float)s_den[0];
  }
}

__device__ __forceinline__ float logaddexp3(float a, float b, float c) {
  const float m = fmaxf(a, fmaxf(b, c));
  if (m == -INFINITY) return -INFINITY;
  return m + logf(expf(a - m) + expf(b - m) + expf(c - m));
}

// One CTA per utterance: alpha recursion over the blank-extended label sequence (S = 2L+1 states) in shared
// memory; log-softmax of each frame is taken on the fly.  Thread s owns state s.
__global__ void ctc_fwd_kernel(const float* __restrict__ logits, const int* __restrict__ labels,
                               const int* __restrict__ label_len, const int* __restrict__ logit_len, int T, int C,
                               int Lmax, int blank, float* __restrict__ loss) {
  extern __shared__ float ctc_smem[];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const int L = min(label_len[b], Lmax), Tb = min(logit_len[b], T);
  const int S = 2 * L + 1;
  float* a0 = ctc_smem;            // [S]
  float* a1 = a0 + (2 * Lmax + 1); // [S]
  float* s_lse = a1 + (2 * Lmax + 1);  // [1]
  __shared__ float s_part[32];
  if (Tb == 0) {
    if (tid == 0) loss[b] = (L == 0) ? 0.f : INFINITY;
    return;
  }
  const int* lab = labels + (size_t)b * Lmax;
  const int ext = (tid < S) ? ((tid & 1) ? max(0, min(lab[tid >> 1], C - 1)) : blank) : 0;
  const int ext2 = (tid >= 2 && tid < S) ? (((tid - 2) & 1) ? max(0, min(lab[(tid - 2) >> 1], C - 1)) : blank) : blank;
  const bool can_skip = (tid < S) && (tid >= 2) && (ext != blank) && (ext != ext2);
  float* cur = a0;
  float* nxt = a1;
  for (int t = 0; t < Tb; ++t) {
    const float* row = logits + ((size_t)b * T + t) * C;
    // block log-sum-exp of the frame
    float m = -INFINITY;
    for (int v = tid; v < C; v += blockDim.x) m = fmaxf(m, row[v]);
    m = warp_max(m);
    if (lane == 0) s_part[warp] = m;
    __syncthreads();
    m = s_part[0];
    for (int w = 1; w < nw; ++w) m = fmaxf(m, s_part[w]);
    __syncthreads();
    float sum = 0.f;
    for (int v = tid; v < C; v += blockDim.x) sum += expf(row[v] - m);
    sum = warp_sum(sum);
    if (lane == 0) s_part[warp] = sum;
    __syncthreads();
    if (tid == 0) {
      float tot = 0.f;
      for (int w = 0; w < nw; ++w) tot += s_part[w];
      *s_lse = m + logf(tot);
    }
    __syncthreads();
    if (tid < S) {
      const float lp = row[ext] - *s_lse;
      float v;
      if (t == 0) {
        v = (tid < 2) ? lp : -INFINITY;
      } else {
        const float p0 = cur[tid];
        const float p1 = tid >= 1 ? cur[tid - 1] : -INFINITY;
        const float p2 = can_skip ? cur[tid - 2] : -INFINITY;
        v = logaddexp3(p0, p1, p2) + lp;
      }
      nxt[tid] = v;
    }
    __syncthreads();
    float* tmp = cur; cur = nxt; nxt = tmp;
  }
  if (tid == 0) {
    const float e1 = cur[S - 1];
    const float e2 = S >= 2 ? cur[S - 2] : -INFINITY;
    loss[b] = -logaddexp3(e1, e2, -INFINITY);
  }
}

}  // namespace plas

using namespace plas;

extern "C" int plas_seq_ce_fwd(const float* logits, const int32_t* targets, const float* weights, int64_t n_tokens,
                               int32_t V, float* ce_tokens, float* out3, plas_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PLAS_REQUIRE(logits && targets && ce_tokens && out3 && n_tokens > 0 && V > 0, "seq_ce: bad argument");
  const int wpb = 8;
  seq_ce_token_kernel<<<(unsigned)((n_tokens + wpb - 1) / wpb), wpb * 32, 0, stream>>>(logits, targets, n_tokens, V, ce_tokens);
  PLAS_CUDA(cudaGetLastError());
  weighted_mean_kernel<<<1, 256, 0, stream>>>(ce_tokens, weights, n_tokens, out3);
  PLAS_CUDA(cudaGetLastError());
  return PLAS_OK;
}

extern "C" int plas_sigmoid_ce_fwd(const float* logits, const float* labels, const float* weights, int64_t n_tokens,
                                   int32_t n_feat, float* ce_tokens, float* out3, plas_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PLAS_REQUIRE(logits && labels && ce_tokens && out3 && n_tokens > 0 && n_feat > 0, "sigmoid_ce: bad argument");
  const int wpb = 8;
  sigmoid_ce_token_kernel<<<(unsigned)((n_tokens + wpb - 1) / wpb), wpb * 32, 0, stream>>>(logits, labels, n_tokens, n_feat, ce_tokens);
  PLAS_CUDA(cudaGetLastError());
  weighted_mean_kernel<<<1, 256, 0, stream>>>(ce_tokens, weights, n_tokens, out3);
  PLAS_CUDA(cudaGetLastError());
  return PLAS_OK;
}

extern "C" int plas_ctc_fwd(const float* logits, const int32_t* labels, const int32_t* label_len,
                            const int32_t* logit_len, int32_t B, int32_t T, int32_t C, int32_t Lmax, int32_t blank,
                            float* loss, plas_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PLAS_REQUIRE(logits && labels && label_len && logit_len && loss, "ctc: null argument");
  PLAS_REQUIRE(B > 0 && T > 0 && C > 1 && Lmax >= 0 && blank >= 0 && blank < C, "ctc: bad shape");
  const int S = 2 * Lmax + 1;
  PLAS_REQUIRE(S <= 1024, "ctc: label length %d too long (2L+1 <= 1024)", Lmax);
  int threads = ((S + 31) / 32) * 32;
  if (threads < 64) threads = 64;
  const size_t smem = (size_t)(2 * S + 4) * sizeof(float);
  ctc_fwd_kernel<<<B, threads, smem, stream>>>(logits, labels, label_len, logit_len, T, C, Lmax, blank, loss);
  PLAS_CUDA(cudaGetLastError());
  return PLAS_OK;
}
