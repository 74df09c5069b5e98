// K5 (training) -- loss heads with their gradients and the optimiser step of the TRAIN graph, exact fp32:
//   plas_seq_ce_grad      tf.contrib.seq2seq.sequence_loss value + d/dlogits              (model_helper.py:24-30)
//   plas_sigmoid_ce_grad  sequence_loss_sigmoid value + d/dlogits                         (model_helper.py:81-105)
//   plas_ctc_grad         tf.nn.ctc_loss_v2 (blank 0) per-utterance value + d/dlogits     (model_helper.py:355-357)
//   plas_grad_l2_norm     g += l2_scale * w (tf.contrib.layers.l2_regularizer, model_helper.py:411-413) and the
//                         per-tensor gradient norms that tf.clip_by_norm(grad, 2) needs   (model_helper.py:416)
//   plas_clip_scale       g *= clip / max(norm, clip) per tensor
//   plas_adam_step        tf.train.AdamOptimizer update, epsilon-hat form                 (model_helper.py:404,417)
// Reductions run in a fixed order (deterministic).  Parameters, gradients and the Adam moments live in flat fp32
// buffers; `offsets` [n_tensors+1] delimits the tensors (the unit of clip_by_norm).
#include "common.cuh"
#include "../../include/plas.h"

namespace plas {

// den[0] = sum(w) (one block, f64)
__global__ void weight_sum_kernel(const float* __restrict__ w, long long n, float* __restrict__ den) {
  __shared__ double s[256];
  double a = 0.0;
  for (long long i = threadIdx.x; i < n; i += 256) a += w ? (double)w[i] : 1.0;
  s[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) den[0] = (float)s[0];
}

// warp per token: ce and dlogits = gscale * w / (sum(w) + 1e-12) * (softmax - onehot)
__global__ void seq_ce_grad_kernel(const float* __restrict__ logits, const int* __restrict__ targets,
                                   const float* __restrict__ w, const float* __restrict__ den, long long n_tok, int V,
                                   float gscale, float* __restrict__ ce, float* __restrict__ dlogits) {
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
  const int tg = max(0, min(targets[tok], V - 1));
  const float lse = m + logf(s);
  if (lane == 0) ce[tok] = lse - row[tg];
  const float coef = gscale * (w ? w[tok] : 1.f) / (den[0] + 1e-12f);
  float* drow = dlogits + tok * V;
  for (int v = lane; v < V; v += 32) drow[v] = coef * (expf(row[v] - lse) - (v == tg ? 1.f : 0.f));
}

__global__ void sigmoid_ce_grad_kernel(const float* __restrict__ logits, const float* __restrict__ labels,
                                       const float* __restrict__ w, const float* __restrict__ den, long long n_tok, int n,
                                       float gscale, float* __restrict__ ce, float* __restrict__ dlogits) {
  const int lane = threadIdx.x & 31;
  const long long tok = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tok >= n_tok) return;
  const float* x = logits + tok * n;
  const float* z = labels + tok * n;
  const float coef = gscale * (w ? w[tok] : 1.f) / (den[0] + 1e-12f) / (float)n;
  float* dx = dlogits + tok * n;
  float s = 0.f;
  for (int k = lane; k < n; k += 32) {
    s += fmaxf(x[k], 0.f) - x[k] * z[k] + log1pf(expf(-fabsf(x[k])));
    dx[k] = coef * (sigmoidf_acc(x[k]) - z[k]);
  }
  s = warp_sum(s);
  if (lane == 0) ce[tok] = s / (float)n;
}

// out[0] = sum(ce*w) / (sum(w) + 1e-12), out[1] = sum(ce*w), out[2] = sum(w)
// compute_log_probs_loss (model_helper.py:132-146) of the --binf_projection speller: the attention vector is read as
// [log p(feature = 1) | log p(feature = 0)]; the regulariser |p1 + p0 - 1| + relu(log p1) + relu(log p0), averaged over every
// element, pushes it towards normalised log-probabilities.  One warp per row; reg_tok[row] = weight * (row mean); the gradient
// treats the stabilising constant c = -(l1 + l0) / 2 as a constant (tf.stop_gradient), like the reference.
__global__ void log_probs_reg_grad_kernel(const float* __restrict__ att, long long n_tok, int n, float weight,
                                          float* __restrict__ reg_tok, float* __restrict__ datt) {
  const int lane = threadIdx.x & 31;
  const long long tok = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tok >= n_tok) return;
  const float* row = att + tok * 2 * n;
  float* drow = datt + tok * 2 * n;
  const float coef = weight / ((float)n_tok * (float)n);
  float s = 0.f;
  for (int k = lane; k < n; k += 32) {
    const float l1 = row[k], l0 = row[n + k];
    const float c = -(l1 + l0) * 0.5f;
    const float e1 = expf(l1 + c), e0 = expf(l0 + c), ec = expf(c);
    const float dev = (e1 + e0) / ec - 1.f;
    s += fabsf(dev) + fmaxf(l1, 0.f) + fmaxf(l0, 0.f);
    const float sg = dev > 0.f ? 1.f : (dev < 0.f ? -1.f : 0.f);
    drow[k] = coef * (sg * e1 / ec + (l1 > 0.f ? 1.f : 0.f));
    drow[n + k] = coef * (sg * e0 / ec + (l0 > 0.f ? 1.f : 0.f));
  }
  s = warp_sum(s);
  if (lane == 0) reg_tok[tok] = weight * s / (float)n;
}

__global__ void weighted_mean2_kernel(const float* __restrict__ ce, const float* __restrict__ w, long long n_tok,
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

__device__ __forceinline__ float lae3(float a, float b, float c) {
  const float m = fmaxf(a, fmaxf(b, c));
  if (m == -INFINITY) return -INFINITY;
  return m + logf(expf(a - m) + expf(b - m) + expf(c - m));
}

// One CTA per utterance.  Pass 1: alpha recursion, every alpha_t kept in `alphas` [B][T][Smax] (global workspace).
// Pass 2: beta recursion backwards; per frame  dlogits[t][c] = gscale * (softmax_t[c] - exp(lse_{s: ext_s = c}(alpha_t(s) +
// beta_t(s)) - logp_t[c] + nll)).  Frames t >= logit_len get zero gradient.
__global__ void ctc_grad_kernel(const float* __restrict__ logits, const int* __restrict__ labels,
                                const int* __restrict__ label_len, const int* __restrict__ logit_len, int T, int C, int Lmax,
                                int blank, float gscale, float* __restrict__ alphas, float* __restrict__ loss,
                                float* __restrict__ dlogits) {
  extern __shared__ float ctc_smem[];
  const int Smax = 2 * Lmax + 1;
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const int L = min(label_len[b], Lmax), Tb = min(logit_len[b], T);
  const int S = 2 * L + 1;
  float* b0 = ctc_smem;          // [Smax + 2] beta ping (two -inf guard slots at the end)
  float* b1 = b0 + Smax + 2;     // [Smax + 2]
  float* s_ab = b1 + Smax + 2;   // [Smax] alpha + beta of the frame
  int* s_ext = reinterpret_cast<int*>(s_ab + Smax);  // [Smax]
  __shared__ float s_part[32];
  __shared__ float s_lse;
  float* drow_all = dlogits + (size_t)b * T * C;
  for (int i = tid; i < T * C; i += blockDim.x) drow_all[i] = 0.f;
  if (Tb == 0) {
    if (tid == 0) loss[b] = (L == 0) ? 0.f : INFINITY;
    return;
  }
  const int* lab = labels + (size_t)b * Lmax;
  for (int s = tid; s < Smax; s += blockDim.x) s_ext[s] = (s < S && (s & 1)) ? max(0, min(lab[s >> 1], C - 1)) : blank;
  __syncthreads();
  float* al = alphas + (size_t)b * T * Smax;
  auto frame_lse = [&](const float* row) {
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
      s_lse = m + logf(tot);
    }
    __syncthreads();
  };
  // ---- alpha ----
  for (int t = 0; t < Tb; ++t) {
    const float* row = logits + ((size_t)b * T + t) * C;
    frame_lse(row);
    for (int s = tid; s < S; s += blockDim.x) {
      const int e = s_ext[s];
      const float lp = row[e] - s_lse;
      float v;
      if (t == 0) {
        v = (s < 2) ? lp : -INFINITY;
      } else {
        const float* prev = al + (size_t)(t - 1) * Smax;
        const bool skip = s >= 2 && e != blank && e != s_ext[s - 2];
        v = lae3(prev[s], s >= 1 ? prev[s - 1] : -INFINITY, skip ? prev[s - 2] : -INFINITY) + lp;
      }
      al[(size_t)t * Smax + s] = v;
    }
    __syncthreads();
  }
  const float* last = al + (size_t)(Tb - 1) * Smax;
  const float logp = lae3(last[S - 1], S >= 2 ? last[S - 2] : -INFINITY, -INFINITY);
  const float nll = -logp;
  if (tid == 0) loss[b] = nll;
  if (!(nll < INFINITY)) return;  // infeasible alignment: loss is inf, gradient left at zero
  // ---- beta + gradient ----
  float* cur = b0;
  float* nxt = b1;
  for (int s = tid; s < Smax + 2; s += blockDim.x) { b0[s] = -INFINITY; b1[s] = -INFINITY; }
  __syncthreads();
  for (int t = Tb - 1; t >= 0; --t) {
    const float* row = logits + ((size_t)b * T + t) * C;
    frame_lse(row);
    for (int s = tid; s < S; s += blockDim.x) {
      const int e = s_ext[s];
      const float lp = row[e] - s_lse;
      float v;
      if (t == Tb - 1) {
        v = (s >= S - 2) ? lp : -INFINITY;
      } else {
        const bool skip = s + 2 < S && s_ext[s + 2] != blank && s_ext[s + 2] != e;
        v = lae3(cur[s], s + 1 < S ? cur[s + 1] : -INFINITY, skip ? cur[s + 2] : -INFINITY) + lp;
      }
      nxt[s] = v;
      s_ab[s] = al[(size_t)t * Smax + s] + v;
    }
    __syncthreads();
    float* drow = drow_all + (size_t)t * C;
    for (int c = tid; c < C; c += blockDim.x) {
      float m = -INFINITY;
      for (int s = 0; s < S; ++s)
        if (s_ext[s] == c) m = fmaxf(m, s_ab[s]);
      float occ = 0.f;
      if (m > -INFINITY) {
        float sum = 0.f;
        for (int s = 0; s < S; ++s)
          if (s_ext[s] == c) sum += expf(s_ab[s] - m);
        const float lpc = row[c] - s_lse;
        occ = expf(m + logf(sum) - lpc + nll);
      }
      drow[c] = gscale * (expf(row[c] - s_lse) - occ);
    }
    __syncthreads();
    float* tmp = cur; cur = nxt; nxt = tmp;
  }
}

// ---- optimiser ------------------------------------------------------------------------------------------
constexpr int L2_SPLIT = 16;  // slices per tensor (stage 1), summed in a fixed order by stage 2

// stage 1: CTA (slice j, tensor i): g += l2*w over its slice, partial sums of g^2 and w^2 -> part[i][j][2]
__global__ void __launch_bounds__(1024) grad_l2_norm_kernel(const float* __restrict__ params, float* __restrict__ grads,
                                                            const long long* __restrict__ offsets, float l2,
                                                            double* __restrict__ part) {
  __shared__ double s[1024];
  __shared__ double s2[1024];
  const long long t_lo = offsets[blockIdx.y], t_hi = offsets[blockIdx.y + 1];
  const long long per = (t_hi - t_lo + L2_SPLIT - 1) / L2_SPLIT;
  const long long lo = t_lo + (long long)blockIdx.x * per;
  const long long hi = lo + per < t_hi ? lo + per : t_hi;
  double a = 0.0, w2 = 0.0;
  for (long long i = lo + threadIdx.x; i < hi; i += 1024) {
    const float w = params[i];
    const float g = fmaf(l2, w, grads[i]);
    grads[i] = g;
    a += (double)g * (double)g;
    w2 += (double)w * (double)w;
  }
  s[threadIdx.x] = a;
  s2[threadIdx.x] = w2;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      s[threadIdx.x] += s[threadIdx.x + o];
      s2[threadIdx.x] += s2[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    part[((size_t)blockIdx.y * L2_SPLIT + blockIdx.x) * 2] = s[0];
    part[((size_t)blockIdx.y * L2_SPLIT + blockIdx.x) * 2 + 1] = s2[0];
  }
}

// stage 2: one thread per tensor, fixed order
__global__ void grad_l2_norm_finish_kernel(const double* __restrict__ part, int n, float* __restrict__ norms, float* __restrict__ wsq) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a = 0.0, w2 = 0.0;
  for (int j = 0; j < L2_SPLIT; ++j) {
    a += part[((size_t)i * L2_SPLIT + j) * 2];
    w2 += part[((size_t)i * L2_SPLIT + j) * 2 + 1];
  }
  norms[i] = (float)sqrt(a);
  if (wsq) wsq[i] = (float)w2;
}

__global__ void clip_scale_kernel(float* __restrict__ grads, const long long* __restrict__ offsets,
                                  const float* __restrict__ norms, float clip, float post) {
  const long long lo = offsets[blockIdx.y], hi = offsets[blockIdx.y + 1];
  const float sc = clip / fmaxf(norms[blockIdx.y], clip) * post;
  for (long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (long long)gridDim.x * blockDim.x)
    grads[i] *= sc;
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr_t, float b1, float b2, float eps, float gscale) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

// y[i] = x[i] * mask(seed, i) / keep  (forward on activations, backward on their gradients: same op, same seed)
__global__ void dropout_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, unsigned seed0,
                               const unsigned* __restrict__ step_ptr, unsigned thresh, float inv_keep) {
  const unsigned seed = seed0 + (step_ptr ? *step_ptr : 0u) * DROP_STEP_MUL;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = x[i] * drop_scale((uint64_t)i, seed, thresh, inv_keep);
}

// periodic Gaussian weight noise (model_helper.py:418-432): x += mean + std * N(0,1), deviates from the counter hash
__global__ void add_normal_noise_kernel(float* __restrict__ x, long long n, unsigned seed0, const unsigned* __restrict__ step,
                                        float mean, float std) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned seed = seed0 + (step ? *step : 0u) * DROP_STEP_MUL;
  x[i] += mean + std * hash_normal((uint64_t)i, seed);
}

__global__ void axpy_kernel(float* __restrict__ y, const float* __restrict__ x, long long n, float alpha) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = fmaf(alpha, x[i], y[i]);
}

}  // namespace plas

using namespace plas;

extern "C" int plas_dropout_f32(const float* x, float* y, int64_t n, uint32_t seed, const uint32_t* step, float keep_prob,
                                plas_stream_t stream_) {
  PLAS_REQUIRE(x && y && n > 0 && keep_prob > 0.f && keep_prob <= 1.f, "dropout: bad argument");
  const int blocks = (int)((n + 1023) / 1024 < 148 * 8 ? (n + 1023) / 1024 : 148 * 8);
  dropout_kernel<<<blocks, 256, 0, (cudaStream_t)stream_>>>(x, y, n, seed, step, (unsigned)(keep_prob * 16777216.0f), 1.0f / keep_prob);
  PLAS_CUDA(cudaGetLastError());
  return PLAS_OK;
}

extern "C" int plas_add_normal_noise_f32(float* x, int64_t n, uint32_t seed, const uint32_t* step, float mean, float std,
                                         plas_stream_t stream_) {
  PLAS_REQUIRE(x != nullptr && n >= 0 && std >= 0.f, "add_normal_noise: bad argument");
  if (n) add_normal_noise_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(x, n, seed, step, mean, std);
  PLAS_CUDA(cudaGetLastError());
  return PLAS_OK;
}

extern "C" int plas_axpy_f32(float* y, const float* x, int64_t n, float alpha, plas_stream_t stream_) {
  PLAS_REQUIRE(y && x && n > 0, "axpy: bad argument");
  const int blocks = (int)((n + 1023) / 1024 < 148 * 8 ? (n + 1023) / 1024 : 148 * 8);
  axpy_kernel<<<blocks, 256, 0, (cudaStream_t)stream_>>>(y, x, n, alpha);
  PLAS_CUDA(cudaGetLastError());
  return PLAS_OK;
}

extern "C" int plas_seq_ce_grad(const float* logits, const int32_t* targets, const float* weights, int64_t n_tokens,
                                int32_t V, float gscale, float* ce_tokens, float* out3, float* dlogits, plas_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  PLAS_REQUIRE(logits && targets && ce_tokens && out3 && dlogits && n_tokens > 0 && V > 0, "seq_ce_grad: bad argument");
  weight_sum_kernel<<<1, 256, 0, st>>>(weights, n_tokens, out3 + 2);
  seq_ce_grad_kernel<<<(unsigned)((n_tokens + 7) / 8), 256, 0, st>>>(logits, targets, weights, out3 + 2, n_tokens, V, gscale,
                                                                     ce_tokens, dlogits);
  weighted_mean2_kernel<<<1, 256, 0, st>>>(ce_tokens, weights, n_tokens, out3);
  PLAS_CUDA(cudaGetLastError());
  return PLAS_OK;
}

extern "C" int plas_sigmoid_ce_grad(const float* logits, const float* labels, const float* weights, int64_t n_tokens,
                                    int32_t n_feat, float gscale, float* ce_tokens, float* out3, float* dlogits,
                                    plas_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  PLAS_REQUIRE(logits && labels && ce_tokens && out3 && dlogits && n_tokens > 0 && n_feat > 0, "sigmoid_ce_grad: bad argument");
  weight_sum_kernel<<<1, 256, 0, st>>>(weights, n_tokens, out3 + 2);
  sigmoid_ce_grad_kernel<<<(unsigned)((n_tokens + 7) / 8), 256, 0, st>>>(logits, labels, weights, out3 + 2, n_tokens, n_feat,
                                                                         gscale, ce_tokens, dlogits);
  weighted_mean2_kernel<<<1, 256, 0, st>>>(ce_tokens, weights, n_tokens, out3);
  PLAS_CUDA(cudaGetLastError());
  return PLAS_OK;
}

extern "C" int plas_log_probs_reg_grad(const float* att, int64_t n_rows, int32_t n_feat, float weight, float* reg_rows, float* out3,
                                       float* datt, plas_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  PLAS_REQUIRE(att && reg_rows && out3 && datt && n_rows > 0 && n_feat > 0, "log_probs_reg_grad: bad argument");
  log_probs_reg_grad_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, st>>>(att, n_rows, n_feat, weight, reg_rows, datt);
  weighted_mean2_kernel<<<1, 256, 0, st>>>(reg_rows, nullptr, n_rows, out3);
  PLAS_CUDA(cudaGetLastError());
  return PLAS_OK;
}

extern "C" size_t plas_ctc_grad_workspace_bytes(int32_t B, int32_t T, int32_t Lmax) {
  return (size_t)B * T * (2 * Lmax + 1) * sizeof(float);
}

extern "C" int plas_ctc_grad(const float* logits, const int32_t* labels, const int32_t* label_len, const int32_t* logit_len,
                             int32_t B, int32_t T, int32_t C, int32_t Lmax, int32_t blank, float gscale, float* loss,
                             float* dlogits, void* workspace, size_t workspace_bytes, plas_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  PLAS_REQUIRE(logits && labels && label_len && logit_len && loss && dlogits && workspace, "ctc_grad: null argument");
  PLAS_REQUIRE(B > 0 && T > 0 && C > 1 && Lmax >= 0 && blank >= 0 && blank < C, "ctc_grad: bad shape");
  PLAS_REQUIRE(workspace_bytes >= plas_ctc_grad_workspace_bytes(B, T, Lmax), "ctc_grad: workspace too small");
  const int S = 2 * Lmax + 1;
  const size_t smem = (size_t)(2 * (S + 2) + 2 * S) * sizeof(float);
  PLAS_REQUIRE(smem <= 200 * 1024, "ctc_grad: label length %d too long", Lmax);
  if (smem > 48 * 1024) PLAS_CUDA(cudaFuncSetAttribute(ctc_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ctc_grad_kernel<<<B, 128, smem, st>>>(logits, labels, label_len, logit_len, T, C, Lmax, blank, gscale, (float*)workspace, loss,
                                        dlogits);
  PLAS_CUDA(cudaGetLastError());
  return PLAS_OK;
}

extern "C" size_t plas_grad_l2_norm_scratch_bytes(int32_t n_tensors) { return (size_t)n_tensors * L2_SPLIT * 2 * sizeof(double); }

extern "C" int plas_grad_l2_norm(const float* params, float* grads, const int64_t* offsets, int32_t n_tensors, float l2_scale,
                                 float* norms, float* wsq, void* scratch, size_t scratch_bytes, plas_stream_t stream_) {
  PLAS_REQUIRE(params && grads && offsets && norms && scratch && n_tensors > 0, "grad_l2_norm: bad argument");
  PLAS_REQUIRE(scratch_bytes >= plas_grad_l2_norm_scratch_bytes(n_tensors), "grad_l2_norm: scratch too small");
  grad_l2_norm_kernel<<<dim3(L2_SPLIT, n_tensors), 1024, 0, (cudaStream_t)stream_>>>(params, grads, (const long long*)offsets, l2_scale,
                                                                                     (double*)scratch);
  grad_l2_norm_finish_kernel<<<(n_tensors + 127) / 128, 128, 0, (cudaStream_t)stream_>>>((const double*)scratch, n_tensors, norms, wsq);
  PLAS_CUDA(cudaGetLastError());
  return PLAS_OK;
}

extern "C" int plas_clip_scale(float* grads, const int64_t* offsets, int32_t n_tensors, const float* norms, float clip,
                               float post_scale, plas_stream_t stream_) {
  PLAS_REQUIRE(grads && offsets && norms && n_tensors > 0 && clip > 0.f, "clip_scale: bad argument");
  clip_scale_kernel<<<dim3(32, n_tensors), 256, 0, (cudaStream_t)stream_>>>(grads, (const long long*)offsets, norms, clip,
                                                                            post_scale);
  PLAS_CUDA(cudaGetLastError());
  return PLAS_OK;
}

extern "C" int plas_adam_step(float* params, const float* grads, float* m, float* v, int64_t n, float lr_t, float beta1,
                              float beta2, float eps, float grad_scale, plas_stream_t stream_) {
  PLAS_REQUIRE(params && grads && m && v && n > 0, "adam_step: bad argument");
  const int blocks = (int)((n + 1023) / 1024 < 148 * 8 ? (n + 1023) / 1024 : 148 * 8);
  adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream_>>>(params, grads, m, v, n, lr_t, beta1, beta2, eps, grad_scale);
  PLAS_CUDA(cudaGetLastError());
  return PLAS_OK;
}
