// K4 (training) -- teacher-forced attention decoder, forward with saved activations and full backward (BPTT), exact
// fp32, directly on the TF checkpoint layout.  Replaces speller(mode=TRAIN) (reference las/model.py:205-296,344-349:
// attend -> AttentionWrapper(MultiRNNCell(LSTMCell)) + BasicDecoder + TrainingHelper / TrainingSigmoidHelper +
// dynamic_decode, projection DenseBinfDecoder utils/training_helper.py:122-153) and its gradient
// (optimizer.compute_gradients, model_helper.py:415).
//
// Structure.  Everything that does not depend on the previous step is hoisted into plas_gemm_f32_ex calls:
//   forward : keys = memory W_mem;  Z_0 = x_in W_0[0:E] + b_0;  logits = Att W_proj + b_proj
//   backward: dAtt(from logits) = dlogits W_proj^T;  all weight gradients (X^T dZ over the B*S saved rows);  per
//             utterance dkeys = dScore^T H_top, dvalues = Align^T dAtt;  dmemory += dvalues + dkeys W_mem^T
// and only the truly sequential part runs per step as small kernels (launched from the C++ loops below; the whole
// training step is captured in a CUDA graph by the host):
//   dec_cell_fwd   z = pre + [in1;h_{t-1}] W   (lane = batch row, CTA = a few units, warps split the reduction)
//   dec_att_fwd    query -> scores over the memory -> masked softmax -> context      (CTA = utterance x D-slice)
//   dec_att_bwd    dAtt_t -> dalign -> dscore -> dquery (luong) / dpq, dkeys, dv, dquery (bahdanau)
//   dec_cell_bwd   gate derivatives, dz_t written IN PLACE over the saved gates
//   dec_gemv_t     dinp = dz_t W^T   (gradient wrt [attention_{t-1}; h_{t-1}] that the next iteration consumes)
// Saved tensors are batch-major [B][S][width] so the hoisted GEMMs see plain row-major matrices.
//
// Variants carried by the same kernels, selected by descriptor fields (plas.h): the five attention mechanisms of
// las/model.py:153-166 (luong, bahdanau, luong_monotonic, bahdanau_monotonic with TRAIN-mode score noise / hard mode outside
// TRAIN, custom), the AttentionMultiCell wiring (bottom_only) with pass_hidden_state, attention_layer_size, input dropout,
// scheduled sampling, a constant projection + extra attention gradient (--binf_projection), the input gradient
// (embedding_size).  This file also holds the fp32 inference loop built from the same step kernels
// (plas_decoder_infer_f32: greedy, teacher-forced, beam search).
#include <cooperative_groups.h>
#include <float.h>

#include "common.cuh"
#include "../../include/plas.h"

namespace plas {

constexpr int DT_ROWS = 32;  // batch rows per CTA (lane = row)

// ---------------------------------------------------------------------------------------------------------
struct CellFwdArgs {
  int B, Ud;
  const float* pre; long long s_pre;    // [B][4Ud] pre-activations already holding x W + b (or NULL)
  const float* bias;                    // [4Ud] added when pre == NULL
  const float* in1; long long s1; int K1;   // attention_{t-1} (layer 0) or h of the layer below
  const float* in2; long long s2; int K2;   // h_{t-1} of this layer (default wiring) / old attention (bottom_only upper cells)
  const float* in3; long long s3; int K3;   // bottom_only upper cells: h_{t-1} of this layer; K3 = 0 otherwise
  const float* w;                       // rows of the TF kernel multiplying [in1; in2]: [K1+K2][4Ud]
  const float* c_prev; long long s_c;   // NULL at t == 0
  float* z_out; long long s_z;          // activated gates (i, tanh j, f, o), gate-blocked [4][Ud]
  float* c_out; float* h_out; long long s_h;
  float* hprev_next;                    // slot t+1 of the h_{t-1} copy (NULL at the last step)
  const int* skip;                      // inference loop: *skip != 0 -> every utterance has finished, the launch is a no-op
  float* hdrop_out;                     // dropped-out copy of h feeding the layer above (NULL: no dropout / top layer)
  long long idx_base; unsigned seed, thresh; float inv_keep; const unsigned* step_ptr;  // mask index of (b,u) = b*s_h + idx_base + u
  // bottom_only with dropout (DropoutWrapper around every cell of the AttentionMultiCell): the first Kdrop input columns are the
  // cell's dropped-out input; the mask is applied as the row is staged (element b*dm_stride + dm_base + k of the cell's
  // [B][S][Kdrop] mask, seed dm_seed) and the dropped row is kept in xdrop for the weight-gradient GEMM.  Kdrop = 0: none.
  int Kdrop; long long dm_stride, dm_base; unsigned dm_seed; float* xdrop; long long s_xd;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

constexpr int CF_KS = 8;    // K-slices = CTAs per cluster
constexpr int CF_UG = 16;   // units per column group (64 gate columns)

// z = pre + [in1; in2] W for one step, as a [32 rows] x [K] x [4Ud] product cut in BOTH directions: grid = (8 K-slices) x
// (Ud/16 column groups) x (row chunks), the 8 K-slices of a column group forming one thread-block cluster.  A CTA stages
// only its K-slice of the activations (32 x K/8) and of the weights (K/8 x 64) -- the activations cross L2->SM once per
// column group instead of once per 2 units -- multiplies (lane = batch row, warp = 2 units x 4 gates), and sends each
// warp's partial sums to the CTA that owns those 2 units through distributed shared memory; after one cluster barrier
// every CTA sums the 8 partials of its units in a fixed order and applies the gate math.
__global__ void __launch_bounds__(256) dec_cell_fwd_kernel(CellFwdArgs p) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) float cf_smem[];
  if (p.skip && *p.skip) return;              // uniform over the grid: taken before any cluster barrier
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ks = blockIdx.x;                  // K-slice = rank in the cluster
  const int u0 = blockIdx.y * CF_UG;
  const int b0 = blockIdx.z * DT_ROWS;
  const int Ud = p.Ud;
  const int K = p.K1 + p.K2 + p.K3, K4 = K / 4;
  const int per = (K4 + CF_KS - 1) / CF_KS;   // float4 chunks of the reduction per slice
  const int q_lo = ks * per, nq = max(0, min(K4, q_lo + per) - q_lo);
  const int AS = 4 * per + 4;                 // padded activation row stride
  float* s_a = cf_smem;                                   // [32][AS]
  float* s_w = s_a + (size_t)DT_ROWS * AS;                // [4*per][64], column = unit_local*4 + gate
  float* s_red = s_w + (size_t)4 * per * 4 * CF_UG;       // [8 src][32 rows][8] partial sums of my 2 units
  cluster.barrier_arrive();  // every CTA of the cluster must be running before its shared memory is written remotely
  for (int i = threadIdx.x; i < DT_ROWS * nq; i += 256) {
    const int r = i / nq, q = i - r * nq;
    const int b = min(b0 + r, p.B - 1);
    const int k = 4 * (q_lo + q);
    const float* src = k < p.K1 ? p.in1 + (long long)b * p.s1 + k
                       : (k < p.K1 + p.K2 ? p.in2 + (long long)b * p.s2 + (k - p.K1) : p.in3 + (long long)b * p.s3 + (k - p.K1 - p.K2));
    cp_async16(s_a + (size_t)r * AS + 4 * q, src);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int i = threadIdx.x; i < 4 * nq * 16; i += 256) {  // 16 float4 per weight row: 4 gates x 4 groups of 4 units
    const int kk = i >> 4, g = (i >> 2) & 3, j = i & 3;
    const float4 w = __ldg(reinterpret_cast<const float4*>(p.w + (size_t)(4 * q_lo + kk) * 4 * Ud + g * Ud + u0 + 4 * j));
    float* dst = s_w + (size_t)kk * 64 + (4 * j) * 4 + g;
    dst[0] = w.x; dst[4] = w.y; dst[8] = w.z; dst[12] = w.w;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if (p.Kdrop > 0 && 4 * q_lo < p.Kdrop) {
    const unsigned dseed = p.dm_seed + (p.step_ptr ? *p.step_ptr : 0u) * DROP_STEP_MUL;
    for (int i = threadIdx.x; i < DT_ROWS * nq * 4; i += 256) {
      const int r = i / (nq * 4), c = i - r * (nq * 4);
      const int k = 4 * q_lo + c;
      const int b = b0 + r;
      if (k < p.Kdrop && b < p.B) {
        const float v = s_a[(size_t)r * AS + c] * drop_scale((uint64_t)((long long)b * p.dm_stride + p.dm_base + k), dseed, p.thresh, p.inv_keep);
        s_a[(size_t)r * AS + c] = v;
        if (blockIdx.y == 0 && p.xdrop) p.xdrop[(long long)b * p.s_xd + k] = v;
      }
    }
    __syncthreads();
  }
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const float* arow = s_a + (size_t)lane * AS;
  const float* wbase = s_w + 8 * warp;
#pragma unroll 2
  for (int q = 0; q < nq; ++q) {
    const float4 a = *reinterpret_cast<const float4*>(arow + 4 * q);
    const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float4 w0 = *reinterpret_cast<const float4*>(wbase + (size_t)(4 * q + kk) * 64);
      const float4 w1 = *reinterpret_cast<const float4*>(wbase + (size_t)(4 * q + kk) * 64 + 4);
      acc[0] = fmaf(av[kk], w0.x, acc[0]); acc[1] = fmaf(av[kk], w0.y, acc[1]);
      acc[2] = fmaf(av[kk], w0.z, acc[2]); acc[3] = fmaf(av[kk], w0.w, acc[3]);
      acc[4] = fmaf(av[kk], w1.x, acc[4]); acc[5] = fmaf(av[kk], w1.y, acc[5]);
      acc[6] = fmaf(av[kk], w1.z, acc[6]); acc[7] = fmaf(av[kk], w1.w, acc[7]);
    }
  }
  cluster.barrier_wait();
  {
    float* remote = cluster.map_shared_rank(s_red, warp) + ((size_t)ks * DT_ROWS + lane) * 8;  // warp w's units belong to rank w
    *reinterpret_cast<float4*>(remote) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(remote + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
  cluster.sync();
  if (threadIdx.x < DT_ROWS * 2) {
    const int r = threadIdx.x % DT_ROWS, uu = threadIdx.x / DT_ROWS;
    const int bb = b0 + r, u = u0 + 2 * ks + uu;
    if (bb < p.B) {
      float z[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) z[g] = p.pre ? p.pre[(long long)bb * p.s_pre + g * Ud + u] : p.bias[g * Ud + u];
#pragma unroll
      for (int src = 0; src < CF_KS; ++src) {  // fixed order: deterministic
        const float4 v = *reinterpret_cast<const float4*>(s_red + ((size_t)src * DT_ROWS + r) * 8 + 4 * uu);
        z[0] += v.x; z[1] += v.y; z[2] += v.z; z[3] += v.w;
      }
      const float cp = p.c_prev ? p.c_prev[(long long)bb * p.s_c + u] : 0.f;
      const float gi = sigmoidf_acc(z[0]), gj = tanhf(z[1]), gf = sigmoidf_acc(z[2] + 1.0f), go = sigmoidf_acc(z[3]);
      const float c = gf * cp + gi * gj;
      const float h = go * tanhf(c);
      float* zo = p.z_out + (long long)bb * p.s_z + u;
      zo[0] = gi; zo[Ud] = gj; zo[2 * Ud] = gf; zo[3 * Ud] = go;
      p.c_out[(long long)bb * p.s_h + u] = c;
      p.h_out[(long long)bb * p.s_h + u] = h;
      if (p.hprev_next) p.hprev_next[(long long)bb * p.s_h + u] = h;
      if (p.hdrop_out)
        p.hdrop_out[(long long)bb * p.s_h + u] =
            h * drop_scale((uint64_t)((long long)bb * p.s_h + p.idx_base + u), p.seed + (p.step_ptr ? *p.step_ptr : 0u) * DROP_STEP_MUL,
                           p.thresh, p.inv_keep);
    }
  }
}

// CustomAttention: keys = relu(memory_layer(values)) and its backward (gradient passes where the output is positive)
__global__ void dec_relu_kernel(float* x, size_t n) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) x[i] = fmaxf(x[i], 0.f);
}
__global__ void dec_relu_grad_kernel(float* dx, const float* y, size_t n) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n && !(y[i] > 0.f)) dx[i] = 0.f;
}

// ---------------------------------------------------------------------------------------------------------
struct AttFwdArgs {
  int B, Tm, D, Ud, type, dsplit, staged;
  const float* keys;     // [B][Tm][Ud]
  const float* values;   // [B][Tm][D]
  const int* mem_len;
  const float* query; long long s_q;      // h_top of this step
  const float* w_query; const float* v_att;  // bahdanau
  float* pq; long long s_pq;              // bahdanau: saved processed query
  float* align; long long s_al;           // [Tm] per row
  float* att; long long s_att;            // context of this step
  float* att_next;                        // slot t+1 of the attention_{t-1} copy (NULL at the last step)
  const int* skip;                        // inference loop: *skip != 0 -> no-op
  // luong_monotonic (tf.contrib.seq2seq.LuongMonotonicAttention): p = sigmoid(score + bias), a = p * cumprod_excl(1-p) * cumsum(a_prev / cumprod)
  const float* score_bias;                // [1] attention_score_bias
  const float* align_prev; long long s_ap;  // alignments of step t-1 (NULL at t == 0: a dirac at frame 0)
  float* p_save; long long s_ps;          // training: the choose probabilities of this step (NULL in inference)
  // bahdanau_monotonic (las/model.py:159-164): TRAIN adds noise_scale * N(0,1) to the scores (sigmoid_noise), otherwise mode 'hard'
  int hard; float noise_scale; unsigned noise_seed; long long noise_base;  // noise index b*s_al + noise_base + t
  long long next_base; unsigned seed, thresh; float inv_keep; const unsigned* step_ptr;  // mask index b*s_att + next_base + d
};

// CTA = (utterance, D-slice).  The utterance's keys and the CTA's slice of the values are staged in shared memory
// with one wave of cp.async copies when they fit (staged != 0), so the step is one L2 round trip deep instead of
// one per memory row; otherwise rows stream from L2.
__global__ void __launch_bounds__(256) dec_att_fwd_kernel(AttFwdArgs p) {
  extern __shared__ __align__(16) float att_smem[];
  if (p.skip && *p.skip) return;
  const int Ud = p.Ud, Tm = p.Tm, D = p.D;
  float* s_q = att_smem;          // [Ud] query (luong) or processed query (bahdanau)
  float* s_v = s_q + Ud;          // [Ud] attention_v (bahdanau)
  float* s_sc = s_v + Ud;         // [Tm rounded to 4]
  float* s_keys = s_sc + ((Tm + 3) & ~3);  // staged: [len][Ud]
  __shared__ float s_red[8];
  __shared__ float s_bcast;
  const int b = blockIdx.x, part = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int len = min(p.mem_len[b], Tm);
  const int dper = (D + p.dsplit - 1) / p.dsplit;
  const int d_lo = part * dper, d_hi = min(D, d_lo + dper);
  const int dw = d_hi - d_lo;
  float* s_vals = s_keys + (size_t)Tm * Ud;  // staged: [len][dper]
  const float* kb = p.keys + (size_t)b * Tm * Ud;
  const float* vb = p.values + (size_t)b * Tm * D;
  if (p.staged) {
    const int kq = len * Ud / 4;
    for (int i = tid; i < kq; i += 256) cp_async16(s_keys + 4 * i, kb + 4 * i);
    const int dq = dw / 4;
    if ((dq & (dq - 1)) == 0) {  // power-of-two slices: shift / mask instead of a run-time division per element
      const int lg = 31 - __clz(dq);
      for (int i = tid; i < len * dq; i += 256) {
        const int t = i >> lg, q = i & (dq - 1);
        cp_async16(s_vals + (size_t)t * dper + 4 * q, vb + (size_t)t * D + d_lo + 4 * q);
      }
    } else {
      for (int i = tid; i < len * dq; i += 256) {
        const int t = i / dq, q = i - t * dq;
        cp_async16(s_vals + (size_t)t * dper + 4 * q, vb + (size_t)t * D + d_lo + 4 * q);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  const float* q = p.query + (long long)b * p.s_q;
  const bool bah = p.type == PLAS_ATT_BAHDANAU || p.type == PLAS_ATT_BAHDANAU_MONOTONIC;
  const bool mono = p.type == PLAS_ATT_LUONG_MONOTONIC || p.type == PLAS_ATT_BAHDANAU_MONOTONIC;
  const bool custom = p.type == PLAS_ATT_CUSTOM;  // CustomAttention (las/model.py:72-101): query = relu(query_layer(h)), luong score
  if (bah || custom) {
    for (int u = tid; u < Ud; u += 256) s_v[u] = q[u];  // raw query staged where attention_v goes afterwards
    __syncthreads();
    for (int u = tid; u < Ud; u += 256) {
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      int k = 0;
      for (; k + 4 <= Ud; k += 4) {
        s0 = fmaf(s_v[k], __ldg(p.w_query + (size_t)k * Ud + u), s0);
        s1 = fmaf(s_v[k + 1], __ldg(p.w_query + (size_t)(k + 1) * Ud + u), s1);
        s2 = fmaf(s_v[k + 2], __ldg(p.w_query + (size_t)(k + 2) * Ud + u), s2);
        s3 = fmaf(s_v[k + 3], __ldg(p.w_query + (size_t)(k + 3) * Ud + u), s3);
      }
      for (; k < Ud; ++k) s0 = fmaf(s_v[k], __ldg(p.w_query + (size_t)k * Ud + u), s0);
      float s = (s0 + s1) + (s2 + s3);
      if (custom) s = fmaxf(s, 0.f);
      s_q[u] = s;
      if (part == 0) p.pq[(long long)b * p.s_pq + u] = s;
    }
    __syncthreads();
    if (bah)
      for (int u = tid; u < Ud; u += 256) s_v[u] = p.v_att[u];
  } else {
    for (int u = tid; u < Ud; u += 256) s_q[u] = q[u];
  }
  if (p.staged) asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const float* kbase = p.staged ? s_keys : kb;
  for (int t = warp; t < Tm; t += 8) {
    float s = 0.f;
    if (t < len) {
      const float* kr = kbase + (size_t)t * Ud;
      if (bah)
        for (int u = lane; u < Ud; u += 32) s = fmaf(s_v[u], tanhf(kr[u] + s_q[u]), s);
      else
        for (int u = lane; u < Ud; u += 32) s = fmaf(kr[u], s_q[u], s);
      s = warp_sum(s);
    }
    if (lane == 0) s_sc[t] = t < len ? s : -INFINITY;
  }
  __syncthreads();
  if (mono && p.hard) {
    // monotonic_attention(mode='hard'): p = [score > 0] * cumsum(prev), a = p * cumprod_excl(1 - p)
    if (tid == 0) {
      const float bias = *p.score_bias;
      float cum = 0.f, cp = 1.f;
      for (int t = 0; t < Tm; ++t) {
        cum += p.align_prev ? p.align_prev[(long long)b * p.s_ap + t] : (t == 0 ? 1.f : 0.f);
        const float pc = (t < len && s_sc[t] + bias > 0.f) ? cum : 0.f;
        const float a = pc * cp;
        cp *= 1.f - pc;
        s_sc[t] = a;
        if (part == 0) p.align[(long long)b * p.s_al + t] = a;
      }
    }
    __syncthreads();
  } else if (mono) {
    // one thread walks the memory in order (the recurrences are sequential and short; same operation order as the oracle)
    if (tid == 0) {
      const float bias = *p.score_bias;
      const unsigned nseed = p.noise_seed + (p.step_ptr ? *p.step_ptr : 0u) * DROP_STEP_MUL;
      float cs = 0.f, run = 0.f;
      for (int t = 0; t < Tm; ++t) {
        float sc = s_sc[t] + bias;
        if (p.noise_scale != 0.f && t < len) sc += p.noise_scale * hash_normal((uint64_t)((long long)b * p.s_al + p.noise_base + t), nseed);
        const float pc = t < len ? sigmoidf_acc(sc) : 0.f;
        const float cp = expf(cs);                                   // exclusive cumprod of (1 - p) in log space
        const float prev = p.align_prev ? p.align_prev[(long long)b * p.s_ap + t] : (t == 0 ? 1.f : 0.f);
        run += prev / fminf(fmaxf(cp, 1e-10f), 1.f);
        const float a = pc * cp * run;
        cs += logf(fminf(fmaxf(1.f - pc, FLT_MIN), 1.f));
        s_sc[t] = a;
        if (part == 0) {
          p.align[(long long)b * p.s_al + t] = a;
          if (p.p_save) p.p_save[(long long)b * p.s_ps + t] = pc;
        }
      }
    }
    __syncthreads();
  } else {
  float m = -INFINITY;
  for (int t = tid; t < Tm; t += 256) m = fmaxf(m, s_sc[t]);
  m = warp_max(m);
  if (lane == 0) s_red[warp] = m;
  __syncthreads();
  if (tid == 0) {
    float mm = s_red[0];
    for (int w = 1; w < 8; ++w) mm = fmaxf(mm, s_red[w]);
    s_bcast = mm;
  }
  __syncthreads();
  m = s_bcast;
  float e_sum = 0.f;
  for (int t = tid; t < Tm; t += 256) {
    const float e = t < len ? expf(s_sc[t] - m) : 0.f;
    s_sc[t] = e;
    e_sum += e;
  }
  e_sum = warp_sum(e_sum);
  __syncthreads();
  if (lane == 0) s_red[warp] = e_sum;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_red[w];
    s_bcast = 1.0f / t;
  }
  __syncthreads();
  const float inv = s_bcast;
  for (int t = tid; t < Tm; t += 256) {
    const float a = s_sc[t] * inv;
    s_sc[t] = a;
    if (part == 0) p.align[(long long)b * p.s_al + t] = a;
  }
  __syncthreads();
  }
  for (int d = d_lo + tid; d < d_hi; d += 256) {
    float c0 = 0.f, c1 = 0.f;
    if (p.staged) {
      const float* vs = s_vals + (d - d_lo);
      int t = 0;
      for (; t + 2 <= len; t += 2) {
        c0 = fmaf(s_sc[t], vs[(size_t)t * dper], c0);
        c1 = fmaf(s_sc[t + 1], vs[(size_t)(t + 1) * dper], c1);
      }
      if (t < len) c0 = fmaf(s_sc[t], vs[(size_t)t * dper], c0);
    } else {
#pragma unroll 8
      for (int t = 0; t < len; ++t) c0 = fmaf(s_sc[t], vb[(size_t)t * D + d], c0);
    }
    const float c = c0 + c1;
    p.att[(long long)b * p.s_att + d] = c;
    if (p.att_next)
      p.att_next[(long long)b * p.s_att + d] =
          p.inv_keep == 1.f ? c : c * drop_scale((uint64_t)((long long)b * p.s_att + p.next_base + d), p.seed + (p.step_ptr ? *p.step_ptr : 0u) * DROP_STEP_MUL, p.thresh, p.inv_keep);
  }
}

// ---------------------------------------------------------------------------------------------------------
struct AttBwdArgs {
  int B, Tm, D, Ud, type, nsplit, staged;
  const float* keys; const float* values; const int* mem_len;
  const float* align; long long s_al;
  float* dctx; long long s_dc;           // in: dlogits W_proj^T of this step; out: + recurrent part
  const float* datt_next; long long s_dn;  // gradient wrt attention_{t} from step t+1's cell input (NULL at t = S-1)
  const float* extra[4]; long long s_extra[4];  // bottom_only: further sources of the attention gradient (NULL = absent)
  // luong_monotonic: saved choose probabilities, previous alignments, the gradient wrt the alignments of this step coming from
  // step t+1 (the alignments are a recurrent state) and the one this step hands to step t-1; d(attention_score_bias) per utterance
  const float* score_p; long long s_sp; const float* align_prev; long long s_ap;
  const float* dalign_in; float* dalign_out; float* dbias_acc;
  float* dscore; long long s_ds;         // [Tm] saved for the hoisted dkeys GEMM (luong)
  float* dq; long long s_dq;             // [Ud] gradient wrt the query (h_top)
  // bahdanau
  const float* pq; long long s_pq; const float* w_query; const float* v_att;
  float* dpq; long long s_dpq; float* dkeys; float* dv_acc;
  long long next_base; unsigned seed, thresh; float inv_keep; const unsigned* step_ptr;  // mask of attention_t as step t+1's input
};

// Cluster of NS CTAs per utterance (grid (B, NS), cluster (1, NS, 1)).  CTA `part` owns a D-slice of the context
// gradient / values and a Ud-slice of the query gradient / keys, staged in shared memory with one wave of cp.async
// copies when they fit.  dalign needs the whole depth: the partial dot products are exchanged through distributed
// shared memory and summed in a fixed order by every CTA (likewise dpq for the bahdanau query layer).
__global__ void __launch_bounds__(256) dec_att_bwd_kernel(AttBwdArgs p) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) float attb_smem[];
  const int Ud = p.Ud, Tm = p.Tm, D = p.D, NS = p.nsplit;
  const int Tp = (Tm + 3) & ~3;
  const int b = blockIdx.x, part = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int dper = D / NS, uper = Ud / NS;
  const int d_lo = part * dper, u_lo = part * uper;
  float* s_dc = attb_smem;               // [dper]
  float* s_part = s_dc + dper;           // [NS][Tp] partial dalign of every CTA of the cluster
  float* s_da = s_part + (size_t)NS * Tp;  // [Tp] dalign -> dscore
  float* s_mono = s_da + Tp;             // [2][Tp] (luong_monotonic: cumprod, running sum)
  float* s_dpq = s_mono + 2 * Tp;        // [Ud] (bahdanau)
  float* s_vals = s_dpq + Ud;            // staged [len][dper]
  float* s_keys = s_vals + (size_t)Tm * dper;  // staged [len][uper]
  __shared__ float s_red[8];
  __shared__ float s_bcast;
  const int len = min(p.mem_len[b], Tm);
  const float* vb = p.values + (size_t)b * Tm * D;
  const float* kb = p.keys + (size_t)b * Tm * Ud;
  if (p.staged) {
    const int dq = dper / 4, uq = uper / 4;
    if (((dq & (dq - 1)) | (uq & (uq - 1))) == 0) {  // power-of-two slices: shift / mask instead of run-time divisions
      const int lgd = 31 - __clz(dq), lgu = 31 - __clz(uq);
      for (int i = tid; i < len * dq; i += 256) {
        const int t = i >> lgd, q = i & (dq - 1);
        cp_async16(s_vals + (size_t)t * dper + 4 * q, vb + (size_t)t * D + d_lo + 4 * q);
      }
      for (int i = tid; i < len * uq; i += 256) {
        const int t = i >> lgu, q = i & (uq - 1);
        cp_async16(s_keys + (size_t)t * uper + 4 * q, kb + (size_t)t * Ud + u_lo + 4 * q);
      }
    } else {
      for (int i = tid; i < len * dq; i += 256) {
        const int t = i / dq, q = i - t * dq;
        cp_async16(s_vals + (size_t)t * dper + 4 * q, vb + (size_t)t * D + d_lo + 4 * q);
      }
      for (int i = tid; i < len * uq; i += 256) {
        const int t = i / uq, q = i - t * uq;
        cp_async16(s_keys + (size_t)t * uper + 4 * q, kb + (size_t)t * Ud + u_lo + 4 * q);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  if (NS > 1) cluster.barrier_arrive();  // every CTA of the cluster must be running before its shared memory is written remotely
  float* dctx = p.dctx + (long long)b * p.s_dc + d_lo;
  for (int d = tid; d < dper; d += 256) {
    float v = dctx[d];
    if (p.datt_next) {
      float gnext = p.datt_next[(long long)b * p.s_dn + d_lo + d];
      if (p.inv_keep != 1.f) gnext *= drop_scale((uint64_t)((long long)b * p.s_dc + p.next_base + d_lo + d), p.seed + (p.step_ptr ? *p.step_ptr : 0u) * DROP_STEP_MUL, p.thresh, p.inv_keep);
      v += gnext;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (p.extra[e]) v += p.extra[e][(long long)b * p.s_extra[e] + d_lo + d];
    s_dc[d] = v;
    dctx[d] = v;
  }
  if (p.staged) asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if (NS > 1) cluster.barrier_wait();
  const float* al = p.align + (long long)b * p.s_al;
  for (int t = warp; t < Tm; t += 8) {
    float s = 0.f;
    if (t < len) {
      const float* vr = p.staged ? s_vals + (size_t)t * dper : vb + (size_t)t * D + d_lo;
      for (int d = lane; d < dper; d += 32) s = fmaf(s_dc[d], vr[d], s);
      s = warp_sum(s);
    }
    if (lane == 0) {
      for (int r = 0; r < NS; ++r) {
        float* remote = NS > 1 ? cluster.map_shared_rank(s_part, r) : s_part;
        remote[part * Tp + t] = s;
      }
    }
  }
  if (NS > 1) cluster.sync(); else __syncthreads();
  float dot = 0.f;
  for (int t = tid; t < Tm; t += 256) {
    float v = 0.f;
    for (int r = 0; r < NS; ++r) v += s_part[r * Tp + t];
    s_da[t] = v;
    dot = fmaf(al[t], v, dot);
  }
  dot = warp_sum(dot);
  if (lane == 0) s_red[warp] = dot;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_red[w];
    s_bcast = t;
  }
  __syncthreads();
  dot = s_bcast;
  const bool bah = p.type == PLAS_ATT_BAHDANAU || p.type == PLAS_ATT_BAHDANAU_MONOTONIC;
  const bool custom = p.type == PLAS_ATT_CUSTOM;
  if (p.type == PLAS_ATT_LUONG_MONOTONIC || p.type == PLAS_ATT_BAHDANAU_MONOTONIC) {
    // a_i = p_i cp_i Q_i with cp = cumprod_excl(1 - p) (log space, clipped), Q_i = sum_{j<=i} a_prev_j / clip(cp_j): one thread,
    // forward recompute, then two reverse running sums
    if (tid == 0) {
      float* s_cp = s_mono;
      float* s_Q = s_mono + Tp;
      float cs = 0.f, run = 0.f;
      for (int t = 0; t < Tm; ++t) {
        const float pc = p.score_p[(long long)b * p.s_sp + t];
        const float cp = expf(cs);
        const float prev = p.align_prev ? p.align_prev[(long long)b * p.s_ap + t] : (t == 0 ? 1.f : 0.f);
        run += prev / fminf(fmaxf(cp, 1e-10f), 1.f);
        s_cp[t] = cp;
        s_Q[t] = run;
        cs += logf(fminf(fmaxf(1.f - pc, FLT_MIN), 1.f));
      }
      float R = 0.f, E = 0.f, dbias = 0.f;
      for (int t = Tm - 1; t >= 0; --t) {
        const float pc = p.score_p[(long long)b * p.s_sp + t];
        const float prev = p.align_prev ? p.align_prev[(long long)b * p.s_ap + t] : (t == 0 ? 1.f : 0.f);
        const float da = s_da[t] + (p.dalign_in ? p.dalign_in[(size_t)b * Tm + t] : 0.f);
        const float cp = s_cp[t], Q = s_Q[t];
        const float c = fminf(fmaxf(cp, 1e-10f), 1.f);
        R += da * pc * cp;                                    // dL/dQ_i summed over i >= t  = dL/dq_t
        if (part == 0) p.dalign_out[(size_t)b * Tm + t] = R / c;  // gradient wrt the previous alignments
        float dcp = da * pc * Q;
        if (cp >= 1e-10f) dcp -= R * prev / (c * c);
        float dp = da * cp * Q;
        const float dlogx = E;                                // sum_{i > t} dcp_i cp_i
        E += dcp * cp;
        const float om = 1.f - pc;
        if (om >= FLT_MIN) dp -= dlogx / om;
        const float ds = t < len ? dp * pc * om : 0.f;
        s_da[t] = ds;
        dbias += ds;
        if (part == 0) p.dscore[(long long)b * p.s_ds + t] = ds;
      }
      if (part == 0) p.dbias_acc[b] += dbias;
    }
  } else {
    for (int t = tid; t < Tm; t += 256) {
      const float ds = al[t] * (s_da[t] - dot);
      s_da[t] = ds;
      if (part == 0) p.dscore[(long long)b * p.s_ds + t] = ds;
    }
  }
  __syncthreads();
  if (bah) {
    float* dkb = p.dkeys + (size_t)b * Tm * Ud;
    for (int uu = tid; uu < uper; uu += 256) {
      const int u = u_lo + uu;
      const float pqu = p.pq[(long long)b * p.s_pq + u];
      const float vu = p.v_att[u];
      float dpq = 0.f, dv = 0.f;
      for (int t = 0; t < len; ++t) {
        const float kv = p.staged ? s_keys[(size_t)t * uper + uu] : kb[(size_t)t * Ud + u];
        const float e = tanhf(kv + pqu);
        const float dpre = s_da[t] * vu * (1.f - e * e);
        dpq += dpre;
        dkb[(size_t)t * Ud + u] += dpre;
        dv = fmaf(s_da[t], e, dv);
      }
      p.dpq[(long long)b * p.s_dpq + u] = dpq;
      p.dv_acc[(size_t)b * Ud + u] += dv;
      for (int r = 0; r < NS; ++r) {
        float* remote = NS > 1 ? cluster.map_shared_rank(s_dpq, r) : s_dpq;
        remote[u] = dpq;
      }
    }
    if (NS > 1) cluster.sync(); else __syncthreads();
    for (int kk = warp; kk < uper; kk += 8) {  // dq[k] = sum_u dpq[u] W_query[k][u]
      const int k = u_lo + kk;
      float s = 0.f;
      const float* wr = p.w_query + (size_t)k * Ud;
      for (int u = lane; u < Ud; u += 32) s = fmaf(s_dpq[u], __ldg(wr + u), s);
      s = warp_sum(s);
      if (lane == 0) p.dq[(long long)b * p.s_dq + k] = s;
    }
  } else {
    for (int uu = tid; uu < uper; uu += 256) {
      float s0 = 0.f, s1 = 0.f;
      if (p.staged) {
        int t = 0;
        for (; t + 2 <= len; t += 2) {
          s0 = fmaf(s_da[t], s_keys[(size_t)t * uper + uu], s0);
          s1 = fmaf(s_da[t + 1], s_keys[(size_t)(t + 1) * uper + uu], s1);
        }
        if (t < len) s0 = fmaf(s_da[t], s_keys[(size_t)t * uper + uu], s0);
      } else {
#pragma unroll 8
        for (int t = 0; t < len; ++t) s0 = fmaf(s_da[t], kb[(size_t)t * Ud + u_lo + uu], s0);
      }
      if (!custom) {
        p.dq[(long long)b * p.s_dq + u_lo + uu] = s0 + s1;
      } else {  // the score read relu(query_layer(h)): back through the relu, then (below) through the query layer
        const int u = u_lo + uu;
        const float dpq = p.pq[(long long)b * p.s_pq + u] > 0.f ? s0 + s1 : 0.f;
        p.dpq[(long long)b * p.s_dpq + u] = dpq;
        for (int r = 0; r < NS; ++r) {
          float* remote = NS > 1 ? cluster.map_shared_rank(s_dpq, r) : s_dpq;
          remote[u] = dpq;
        }
      }
    }
    if (custom) {
      if (NS > 1) cluster.sync(); else __syncthreads();
      for (int kk = warp; kk < uper; kk += 8) {  // dq[k] = sum_u dpq[u] W_query[k][u]
        const int k = u_lo + kk;
        float s = 0.f;
        const float* wr = p.w_query + (size_t)k * Ud;
        for (int u = lane; u < Ud; u += 32) s = fmaf(s_dpq[u], __ldg(wr + u), s);
        s = warp_sum(s);
        if (lane == 0) p.dq[(long long)b * p.s_dq + k] = s;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
struct CellBwdArgs {
  int B, Ud;
  float* z; long long s_z;               // saved gates of this step -> dz (in place)
  const float* c_t; const float* c_prev; long long s_c;   // c_prev NULL at t == 0 (or the initial state)
  long long s_cp;                        // row stride of c_prev
  float* dc_carry;                       // [B][Ud] in/out
  int first;                             // 1 at t == S-1: the carry is zero
  const float* dq; long long s_dq;       // top layer: gradient wrt the query (NULL otherwise)
  const float* dh_next; long long s_dn;  // h-part of dinp of this layer from step t+1 (NULL at t == S-1)
  const float* dh_above; long long s_da; // in1-part of dinp of the layer above, same step (NULL for the top layer)
  const float* dh_extra; long long s_dx; // a further, never dropped-out reader of h (the attention layer's h part) or NULL
  long long idx_base; unsigned seed, thresh; float inv_keep; const unsigned* step_ptr;  // mask of this layer's h feeding the layer above
};

__global__ void __launch_bounds__(256) dec_cell_bwd_kernel(CellBwdArgs p) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= p.B * p.Ud) return;
  const int b = i / p.Ud, u = i % p.Ud;
  const int Ud = p.Ud;
  float dh = 0.f;
  if (p.dq) dh += p.dq[(long long)b * p.s_dq + u];
  if (p.dh_next) dh += p.dh_next[(long long)b * p.s_dn + u];
  if (p.dh_extra) dh += p.dh_extra[(long long)b * p.s_dx + u];
  if (p.dh_above) {
    float ga = p.dh_above[(long long)b * p.s_da + u];
    if (p.inv_keep != 1.f) ga *= drop_scale((uint64_t)((long long)b * p.s_c + p.idx_base + u), p.seed + (p.step_ptr ? *p.step_ptr : 0u) * DROP_STEP_MUL, p.thresh, p.inv_keep);
    dh += ga;
  }
  float* zp = p.z + (long long)b * p.s_z + u;
  const float gi = zp[0], gj = zp[Ud], gf = zp[2 * Ud], go = zp[3 * Ud];
  const float c = p.c_t[(long long)b * p.s_c + u];
  const float cp = p.c_prev ? p.c_prev[(long long)b * p.s_cp + u] : 0.f;
  const float tc = tanhf(c);
  const float carry = p.first ? 0.f : p.dc_carry[(size_t)b * Ud + u];
  const float dc = carry + dh * go * (1.f - tc * tc);
  zp[0] = dc * gj * gi * (1.f - gi);
  zp[Ud] = dc * gi * (1.f - gj * gj);
  zp[2 * Ud] = dc * cp * gf * (1.f - gf);
  zp[3 * Ud] = dh * tc * go * (1.f - go);
  p.dc_carry[(size_t)b * Ud + u] = dc * gf;
}

// dst[b][i] = (a ? a[b][i] : 0) + (c ? m[b][i] c[b][i] : 0) on strided rows; m = the input-dropout multiplier of element
// b * s_mask + mask_base + i (1 when inv_keep == 1).  Assembles the gradient wrt the attention vector (a = datt_t, c = cell 0's
// input gradient of step t+1) and, with a = NULL, writes the dropped-out copy of attention_t that step t+1 reads.
__global__ void dec_add_rows_kernel(float* __restrict__ dst, long long s_dst, const float* __restrict__ a, long long s_a,
                                    const float* __restrict__ c, long long s_c, int B, int n, long long s_mask = 0, long long mask_base = 0,
                                    unsigned seed = 0, unsigned thresh = 0, float inv_keep = 1.f, const unsigned* step_ptr = nullptr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * n) return;
  const int b = i / n, k = i - b * n;
  float v = a ? a[(long long)b * s_a + k] : 0.f;
  if (c) {
    float g = c[(long long)b * s_c + k];
    if (inv_keep != 1.f) g *= drop_scale((uint64_t)((long long)b * s_mask + mask_base + k), seed + (step_ptr ? *step_ptr : 0u) * DROP_STEP_MUL, thresh, inv_keep);
    v += g;
  }
  dst[(long long)b * s_dst + k] = v;
}

// dinp[b][k] = sum_n dz[b][n] W[k][n], k in [0, K): lane = batch row, CTA = 8 consecutive k, warps split n
struct GemvTArgs {
  int B, N, K;                           // N = 4Ud
  const float* dz; long long s_z;
  const float* w;                        // [K][N] rows of the TF kernel below the hoisted x rows
  float* dinp; long long s_o;
};

constexpr int GV_NS = 8;    // N-slices = CTAs per cluster
constexpr int GV_KG = 64;   // outputs per CTA group

// dinp[b][k] = sum_n dz[b][n] W[k][n]: grid = (8 N-slices) x (K/64 output groups) x (row chunks), the 8 slices of a group in
// one cluster.  Same scheme as dec_cell_fwd_kernel: stage the slice (32 x N/8 of dz, 64 x N/8 of W), multiply (lane = batch
// row, warp = 8 outputs), send the partials to the owning CTA through distributed shared memory, sum in a fixed order.
__global__ void __launch_bounds__(256) dec_gemv_t_kernel(GemvTArgs p) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) float gv_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ns = blockIdx.x;
  const int k0 = blockIdx.y * GV_KG;
  const int b0 = blockIdx.z * DT_ROWS;
  const int N4 = p.N / 4;
  const int per = (N4 + GV_NS - 1) / GV_NS;
  const int q_lo = ns * per, nq = max(0, min(N4, q_lo + per) - q_lo);
  const int AS = 4 * per + 4;
  float* s_a = gv_smem;                                 // [32][AS] slice of the dz rows
  float* s_w = s_a + (size_t)DT_ROWS * AS;              // [64][4*per] slice of the CTA group's weight rows
  float* s_red = s_w + (size_t)GV_KG * 4 * per;         // [8 src][32 rows][8] partial sums of my 8 outputs
  cluster.barrier_arrive();
  for (int i = threadIdx.x; i < DT_ROWS * nq; i += 256) {
    const int r = i / nq, q = i - r * nq;
    const int b = min(b0 + r, p.B - 1);
    cp_async16(s_a + (size_t)r * AS + 4 * q, p.dz + (long long)b * p.s_z + 4 * (q_lo + q));
  }
  for (int i = threadIdx.x; i < GV_KG * nq; i += 256) {
    const int kk = i / nq, q = i - kk * nq;
    const int k = min(k0 + kk, p.K - 1);
    cp_async16(s_w + (size_t)kk * 4 * per + 4 * q, p.w + (size_t)k * p.N + 4 * (q_lo + q));
  }
  cp_async_wait_all();
  __syncthreads();
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const float* arow = s_a + (size_t)lane * AS;
  const float* wbase = s_w + (size_t)(8 * warp) * 4 * per;
#pragma unroll 2
  for (int q = 0; q < nq; ++q) {
    const float4 a = *reinterpret_cast<const float4*>(arow + 4 * q);
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const float4 w = *reinterpret_cast<const float4*>(wbase + (size_t)kk * 4 * per + 4 * q);
      acc[kk] = fmaf(a.x, w.x, acc[kk]);
      acc[kk] = fmaf(a.y, w.y, acc[kk]);
      acc[kk] = fmaf(a.z, w.z, acc[kk]);
      acc[kk] = fmaf(a.w, w.w, acc[kk]);
    }
  }
  cluster.barrier_wait();
  {
    float* remote = cluster.map_shared_rank(s_red, warp) + ((size_t)ns * DT_ROWS + lane) * 8;  // warp w's outputs belong to rank w
    *reinterpret_cast<float4*>(remote) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(remote + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
  cluster.sync();
  {
    const int r = threadIdx.x % DT_ROWS, j = threadIdx.x / DT_ROWS;  // 256 threads = 32 rows x my 8 outputs
    const int bb = b0 + r, k = k0 + 8 * ns + j;
    if (bb < p.B && k < p.K) {
      float sum = 0.f;
#pragma unroll
      for (int src = 0; src < GV_NS; ++src) sum += s_red[((size_t)src * DT_ROWS + r) * 8 + j];  // fixed order
      p.dinp[(long long)bb * p.s_o + k] = sum;
    }
  }
}

// Scheduled sampling of the phone speller (ScheduledEmbeddingTrainingHelper / TPUScheduledEmbeddingTrainingHelper,
// utils/training_helper.py:48-87, las/model.py:279-288): with probability p per (utterance, step) the NEXT step's input is
// the one-hot of an id drawn from Categorical(logits_t) instead of the target.  One CTA per utterance: rows that are not
// selected exit at once; a selected row computes its logits, draws by Gumbel-max, and rewrites x_in[b][t+1] and the hoisted
// pre-activations Z_0[b][t+1] = W_0[id] + b_0.  Counter-based randomness (same hash as dropout): nothing differentiable
// flows through the draw, so the backward pass is untouched -- it simply sees the inputs that were actually fed.
struct SchedSampleArgs {
  const float* out; long long s_out; int Dout, V;   // what the projection reads at step t
  const float* w_proj; const float* b_proj;
  float* x_next; long long s_x;                     // x_in[.][t+1][:], E == V
  float* z_next; long long s_z; int N;              // Z_0[.][t+1][:], N = 4Ud
  const float* w0; const float* b0;                 // cell-0 kernel rows of the one-hot input [V][N], bias [N]
  long long idx_base; int S;                        // flat (b, t) index = b*S + t (idx_base = t)
  unsigned seed_sel, seed_cat, p_thresh; const unsigned* step_ptr;
  unsigned xdrop_seed, xdrop_thresh; float xdrop_inv_keep;  // input dropout of x_{t+1} (thresh 2^24 / inv_keep 1 = off)
  // embedding_fn other than one-hot (target_embedding, or the binary-feature columns of --binf_projection): the sampled id feeds
  // row `id` of table [V][E]; fed_ids[b][t+1] records it for the input-side gradient.  table == NULL: one-hot, E == V.
  const float* table; int E; int* fed_ids;
};

__global__ void __launch_bounds__(256) dec_sched_sample_kernel(SchedSampleArgs p) {
  extern __shared__ float ss_smem[];
  const int b = blockIdx.x, tid = threadIdx.x;
  const unsigned stepv = (p.step_ptr ? *p.step_ptr : 0u) * DROP_STEP_MUL;
  const uint64_t bt = (uint64_t)b * p.S + p.idx_base;
  if ((drop_hash(bt, p.seed_sel + stepv) >> 8) >= p.p_thresh) return;  // this row keeps its teacher input
  float* s_a = ss_smem;            // [Dout]
  float* s_part = s_a + p.Dout;    // [4][V]
  __shared__ int s_best;
  const int V = p.V, D = p.Dout;
  for (int d = tid; d < D; d += 256) s_a[d] = p.out[(long long)b * p.s_out + d];
  __syncthreads();
  for (int i = tid; i < 4 * V; i += 256) {
    const int sl = i / V, v = i - sl * V;
    const int dper = (D + 3) / 4;
    const int d_hi = min(D, (sl + 1) * dper);
    float acc = 0.f;
    for (int d = sl * dper; d < d_hi; ++d) acc = fmaf(s_a[d], __ldg(p.w_proj + (size_t)d * V + v), acc);
    s_part[i] = acc;
  }
  __syncthreads();
  float* s_logit = s_a;
  for (int v = tid; v < V; v += 256) {
    const float lg = ((s_part[v] + s_part[V + v]) + (s_part[2 * V + v] + s_part[3 * V + v])) + p.b_proj[v];
    const float u = ((float)(drop_hash(bt * V + v, p.seed_cat + stepv) >> 8) + 0.5f) * (1.0f / 16777216.0f);
    s_logit[v] = lg - logf(-logf(u));  // Gumbel-max: argmax(logits + g) ~ Categorical(softmax(logits))
  }
  __syncthreads();
  if (tid == 0) {
    int best = 0;
    float bv = s_logit[0];
    for (int v = 1; v < V; ++v)
      if (s_logit[v] > bv) { bv = s_logit[v]; best = v; }
    s_best = best;
  }
  __syncthreads();
  const int best = s_best;
  if (p.fed_ids && tid == 0) p.fed_ids[(long long)b * p.S + p.idx_base + 1] = best;
  if (p.table) {  // dense input row: x_{t+1} = table[best] (through its dropout mask), Z_0 = x W_0[0:E] + b_0
    float* s_x = ss_smem;  // [E] (the projection scratch is free now)
    __syncthreads();
    for (int e = tid; e < p.E; e += 256) {
      const float sc = p.xdrop_inv_keep == 1.f ? 1.f
                                               : drop_scale((uint64_t)((long long)b * p.s_x + (long long)(p.idx_base + 1) * p.E + e),
                                                            p.xdrop_seed + stepv, p.xdrop_thresh, p.xdrop_inv_keep);
      const float xv = p.table[(size_t)best * p.E + e] * sc;
      s_x[e] = xv;
      p.x_next[(long long)b * p.s_x + e] = xv;
    }
    __syncthreads();
    for (int n = tid; n < p.N; n += 256) {
      float acc = p.b0[n];
      for (int e = 0; e < p.E; ++e) acc = fmaf(s_x[e], p.w0[(size_t)e * p.N + n], acc);
      p.z_next[(long long)b * p.s_z + n] = acc;
    }
    return;
  }
  const float sc = p.xdrop_inv_keep == 1.f ? 1.f
                                           : drop_scale((uint64_t)((long long)b * p.s_x + (p.idx_base + 1) * V + best), p.xdrop_seed + stepv,
                                                        p.xdrop_thresh, p.xdrop_inv_keep);
  for (int v = tid; v < V; v += 256) p.x_next[(long long)b * p.s_x + v] = v == best ? sc : 0.f;
  for (int n = tid; n < p.N; n += 256) p.z_next[(long long)b * p.s_z + n] = fmaf(sc, p.w0[(size_t)best * p.N + n], p.b0[n]);
}

// ---------------------------------------------------------------------------------------------------------
// fp32 inference (reference-precision mode) built from the same step kernels: greedy (GreedyEmbeddingHelper +
// dynamic_decode, las/model.py:337-347) or teacher-forced decoding as a host loop of small launches over a 2-slot state
// ring, replacing the persistent SIMT kernel of decoder.cu for luong / bahdanau attention (240 us -> ~45 us per step at c1).
// ---------------------------------------------------------------------------------------------------------
struct InferState {
  int* cur_ids;    // [B] input ids of the next step
  int* finished;   // [B]
  int* done;       // [1] every utterance has finished (or max_iter reached)
  int* max_iter;   // [1]
};

__global__ void dec_infer_init_kernel(InferState st, const int* __restrict__ mem_len, int B, int max_steps, int teacher_forced,
                                      float factor, int sos_id, int* __restrict__ seq_len, int* __restrict__ n_steps) {
  __shared__ int s_max;
  if (threadIdx.x == 0) {
    int mi = max_steps;
    if (!teacher_forced) {  // greedy stop: rint(max(mem_len) * factor)  (las/model.py:270-274)
      int ml = 0;
      for (int b = 0; b < B; ++b) ml = max(ml, mem_len[b]);
      mi = min(mi, (int)rintf((float)ml * factor));
    }
    s_max = mi;
    *st.max_iter = mi;
    *st.done = mi <= 0 ? 1 : 0;
    *n_steps = 0;
  }
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    st.cur_ids[b] = sos_id;
    st.finished[b] = s_max <= 0 ? 1 : 0;
    seq_len[b] = 0;
  }
}

// pre[b][:] = W_0[id_b][:] + b_0: the one-hot input picks a row of the cell-0 kernel (embedding_fn, las/model.py:245-246)
__global__ void dec_infer_embed_kernel(InferState st, const int* __restrict__ forced, long long s_forced, const float* __restrict__ w0,
                                       const float* __restrict__ b0, int B, int N, int V, float* __restrict__ pre, long long s_pre) {
  if (*st.done) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * N) return;
  const int b = i / N, n = i - b * N;
  int id = forced ? forced[(long long)b * s_forced] : st.cur_ids[b];
  id = max(0, min(id, V - 1));
  pre[(long long)b * s_pre + n] = w0[(size_t)id * N + n] + b0[n];
}

// logits = context W_proj + b (DenseBinfDecoder), greedy argmax (lowest index wins ties); one CTA per utterance
__global__ void __launch_bounds__(256) dec_infer_sample_kernel(InferState st, const float* __restrict__ att, long long s_att, int D, int V,
                                                               const float* __restrict__ w_proj, const float* __restrict__ b_proj,
                                                               float* __restrict__ logits, long long s_logits,
                                                               int* __restrict__ sample_ids, long long s_ids) {
  extern __shared__ float sm_smem[];
  if (*st.done) return;
  float* s_a = sm_smem;       // [D]
  float* s_part = s_a + D;    // [4][V]
  const int b = blockIdx.x, tid = threadIdx.x;
  for (int d = tid; d < D; d += 256) s_a[d] = att[(long long)b * s_att + d];
  __syncthreads();
  const int slices = 4;
  for (int i = tid; i < slices * V; i += 256) {
    const int sl = i / V, v = i - sl * V;
    const int dper = (D + slices - 1) / slices;
    const int d_hi = min(D, (sl + 1) * dper);
    float acc = 0.f;
    for (int d = sl * dper; d < d_hi; ++d) acc = fmaf(s_a[d], __ldg(w_proj + (size_t)d * V + v), acc);
    s_part[i] = acc;
  }
  __syncthreads();
  float* s_logit = s_a;  // reuse
  for (int v = tid; v < V; v += 256) {
    const float lg = ((s_part[v] + s_part[V + v]) + (s_part[2 * V + v] + s_part[3 * V + v])) + b_proj[v];
    s_logit[v] = lg;
    logits[(long long)b * s_logits + v] = lg;
  }
  __syncthreads();
  if (tid == 0) {
    int best = 0;
    float bv = s_logit[0];
    for (int v = 1; v < V; ++v)
      if (s_logit[v] > bv) { bv = s_logit[v]; best = v; }
    sample_ids[(long long)b * s_ids] = best;
    st.cur_ids[b] = best;
  }
}

// finished / sequence-length bookkeeping of dynamic_decode (one thread: B is small and the order must be fixed)
__global__ void dec_infer_finish_kernel(InferState st, int B, int t, int eos_id, int teacher_forced, int* __restrict__ seq_len,
                                        int* __restrict__ n_steps) {
  if (threadIdx.x != 0 || *st.done) return;
  const int max_iter = *st.max_iter;
  int all = 1;
  for (int b = 0; b < B; ++b) {
    const int was = st.finished[b];
    if (!was) seq_len[b] = t + 1;
    const int now = was || (st.cur_ids[b] == eos_id) || (t + 1 >= max_iter);
    st.finished[b] = now;
    all &= now;
  }
  *n_steps = t + 1;
  if ((all && !teacher_forced) || t + 1 >= max_iter) *st.done = 1;  // teacher forcing runs exactly max_steps
}

// ---------------------------------------------------------------------------------------------------------
// Beam search (tf.contrib.seq2seq.BeamSearchDecoder with length_penalty_weight = 0, las/model.py:219-226,298-319) on the same
// step kernels: the batch is the tiled one (row b*W + w = beam w of utterance b).  Per step the logits of the W beams of an
// utterance are turned into W*V candidate scores, the best W are kept (ties: the lower flat index, like tf.nn.top_k), and
// every piece of decoder state is gathered by parent beam.
// ---------------------------------------------------------------------------------------------------------
struct BeamState {
  float* logp;    // [B] total log-probability of each live hypothesis
  int* fin;       // [B] (InferState.finished)
  int* len;       // [B] BeamSearchDecoderState.lengths
  int* parent;    // [S][B] parent beam of each step
  int* word;      // [S][B] word id of each step
};

__global__ void dec_beam_init_kernel(InferState st, BeamState bs, int B, int W, int sos_id, int* __restrict__ seq_len) {
  const int max_iter = *st.max_iter;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const int w = i % W;
    bs.logp[i] = w == 0 ? 0.f : -INFINITY;   // initialize(): one_hot(0, W, on_value 0, off_value -inf)
    st.finished[i] = (w != 0 || max_iter <= 0) ? 1 : 0;  // BeamSearchDecoder._finished: one_hot(0, W, on False, off True)
    bs.len[i] = 0;
    st.cur_ids[i] = sos_id;
    seq_len[i] = 0;
  }
}

// one CTA per utterance: scores of the W*V continuations, W rounds of block arg-max, bookkeeping of _beam_search_step
__global__ void __launch_bounds__(256) dec_beam_step_kernel(InferState st, BeamState bs, const float* __restrict__ logits, int B, int W, int V,
                                                            int t, int eos_id, int* __restrict__ seq_len) {
  extern __shared__ float bm_smem[];
  if (*st.done) return;
  float* s_tot = bm_smem;                       // [W][V]
  __shared__ float s_val[8];
  __shared__ int s_idx[8];
  __shared__ int s_chosen[32];
  __shared__ float s_cval[32];
  __shared__ float s_lse[32];
  __shared__ int s_fin[32], s_len[32];
  __shared__ float s_lp[32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* lg = logits + (size_t)b * W * V;
  if (tid < W) { s_fin[tid] = st.finished[b * W + tid]; s_len[tid] = bs.len[b * W + tid]; s_lp[tid] = bs.logp[b * W + tid]; }
  for (int w = warp; w < W; w += 8) {           // log-sum-exp of each beam's logits
    float m = -INFINITY;
    for (int v = lane; v < V; v += 32) m = fmaxf(m, lg[(size_t)w * V + v]);
    m = warp_max(m);
    float s = 0.f;
    for (int v = lane; v < V; v += 32) s += expf(lg[(size_t)w * V + v] - m);
    s = warp_sum(s);
    if (lane == 0) s_lse[w] = m + logf(s);
  }
  __syncthreads();
  for (int i = tid; i < W * V; i += 256) {
    const int w = i / V, v = i - w * V;
    // _mask_probs: a finished beam is continued by eos only, at no cost (every other word gets the most negative float)
    const float step = s_fin[w] ? (v == eos_id ? 0.f : -FLT_MAX) : lg[i] - s_lse[w];
    s_tot[i] = s_lp[w] + step;
  }
  __syncthreads();
  for (int k = 0; k < W; ++k) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = tid; i < W * V; i += 256) {
      bool taken = false;
      for (int j = 0; j < k; ++j) taken |= s_chosen[j] == i;
      const float v = s_tot[i];
      if (!taken && (v > bv || (v == bv && i < bi))) { bv = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { s_val[warp] = bv; s_idx[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int w8 = 1; w8 < 8; ++w8)
        if (s_val[w8] > bv || (s_val[w8] == bv && s_idx[w8] < bi)) { bv = s_val[w8]; bi = s_idx[w8]; }
      s_chosen[k] = bi;
      s_cval[k] = bv;
    }
    __syncthreads();
  }
  if (tid < W) {
    const int i = b * W + tid;
    const int idx = s_chosen[tid];
    const int beam = idx / V, word = idx - beam * V;
    const int prev_fin = s_fin[beam];
    if (!s_fin[tid]) seq_len[i] = t + 1;        // dynamic_decode's sequence_lengths follow the slot's previous finished flag
    bs.logp[i] = s_cval[tid];
    st.finished[i] = prev_fin | (word == eos_id);
    bs.len[i] = s_len[beam] + (prev_fin ? 0 : 1);
    bs.parent[(size_t)t * B + i] = beam;
    bs.word[(size_t)t * B + i] = word;
    st.cur_ids[i] = word;
  }
}

// dst[i][:] = src[(i / W) * W + parent[i]][:]: the cell state of every hypothesis follows its parent beam
__global__ void dec_beam_gather_kernel(InferState st, float* __restrict__ dst, const float* __restrict__ src, long long s_src, int width,
                                       const int* __restrict__ parent, int B, int W) {
  if (*st.done) return;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * width) return;
  const int r = (int)(i / width), c = (int)(i - (long long)r * width);
  dst[i] = src[(long long)((r / W) * W + parent[r]) * s_src + c];
}
__global__ void dec_beam_scatter_kernel(InferState st, float* __restrict__ dst, long long s_dst, const float* __restrict__ src, int width, int B) {
  if (*st.done) return;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * width) return;
  const int r = (int)(i / width), c = (int)(i - (long long)r * width);
  dst[(long long)r * s_dst + c] = src[i];
}

__global__ void dec_beam_finish_kernel(InferState st, int B, int t, int* __restrict__ n_steps) {
  if (threadIdx.x != 0 || *st.done) return;
  int all = 1;
  for (int i = 0; i < B; ++i) all &= st.finished[i];
  *n_steps = t + 1;
  if (all || t + 1 >= *st.max_iter) *st.done = 1;
}

// beam_search_ops.cc gather_tree: back-trace every final hypothesis through its parents; after the first eos only eos
__global__ void dec_beam_gather_tree_kernel(BeamState bs, const int* __restrict__ n_steps, int B, int W, int S, int eos_id,
                                            int* __restrict__ predicted) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const int b = i / W, w = i - b * W;
  const int T = *n_steps;
  int* out = predicted + (size_t)b * S * W + w;  // [b][t][w]
  for (int t = 0; t < S; ++t) out[(size_t)t * W] = eos_id;
  int max_len = 0;
  for (int j = 0; j < W; ++j) max_len = max(max_len, bs.len[b * W + j]);
  max_len = min(max_len, T);
  if (max_len <= 0) return;
  out[(size_t)(max_len - 1) * W] = bs.word[(size_t)(max_len - 1) * B + i];
  int parent = bs.parent[(size_t)(max_len - 1) * B + i];
  for (int level = max_len - 2; level >= 0; --level) {
    out[(size_t)level * W] = bs.word[(size_t)level * B + b * W + parent];
    parent = bs.parent[(size_t)level * B + b * W + parent];
  }
  bool fin = false;
  for (int t = 0; t < max_len; ++t) {
    if (fin) out[(size_t)t * W] = eos_id;
    else if (out[(size_t)t * W] == eos_id) fin = true;
  }
}

// ---------------------------------------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------------------------------------
struct DecTrainWs {
  size_t keys, att, att_prev, align, pq, dscore, dpq, dinp[4], dq, dc[4], dkeys, dv_acc, dctx, gh, ctx, datt, dqx, psave, dalign, dbias, xdrop[4];
  size_t z[4], c[4], h[4], hprev[4], hdrop[4];
  size_t splitk, splitk_bytes;
  size_t total;
};

static DecTrainWs dec_train_ws(const plas_dec_train_desc& d) {
  DecTrainWs w;
  size_t off = 0;
  auto take = [&](size_t n_floats) {
    const size_t o = off;
    off += (n_floats * 4 + 255) & ~size_t(255);
    return o;
  };
  const size_t B = d.B, S = d.S, Tm = d.Tm, D = d.D, Ud = d.Ud;
  w.keys = take(B * Tm * Ud);
  const size_t Aw = d.att_layer > 0 ? (size_t)d.att_layer : D;  // width of the attention vector that is fed back and projected
  w.att = take(B * S * Aw);
  w.att_prev = take(B * S * Aw);
  w.align = take(B * S * Tm);
  w.pq = take(B * S * Ud);
  w.dscore = take(B * S * Tm);
  w.dpq = take(B * S * Ud);
  w.dq = take(B * Ud);
  w.dkeys = take(B * Tm * Ud);
  w.dv_acc = take(B * Ud);
  w.dctx = take(B * S * D);
  w.gh = take(B * S * Ud);  // bottom_only: dlogits W_proj^T when the projection reads the top cell
  w.ctx = w.datt = w.dqx = 0;
  const bool mono = d.attention_type == PLAS_ATT_LUONG_MONOTONIC || d.attention_type == PLAS_ATT_BAHDANAU_MONOTONIC;
  w.psave = take(mono ? B * S * Tm : 0);   // choose probabilities of every step
  w.dalign = take(mono ? 2 * B * Tm : 0);  // gradient wrt the alignment state, two slots (parity of t)
  w.dbias = take(mono ? B : 0);            // d(attention_score_bias) per utterance
  if (d.att_layer > 0) {    // attention_layer_size: the context and the gradient wrt the (A-wide) attention are kept separately
    w.ctx = take(B * S * D);
    w.datt = take(B * S * (size_t)d.att_layer);
    w.dqx = take(B * Ud);
  }
  for (int l = 0; l < 4; ++l) {
    w.z[l] = w.c[l] = w.h[l] = w.hprev[l] = w.dinp[l] = w.dc[l] = w.hdrop[l] = w.xdrop[l] = 0;
    if (l >= d.n_layers) continue;
    if (d.bottom_only && d.keep_prob < 1.f)  // dropped-out inputs of cell l: attention_{t-1} (l = 0) or [output below; old attention]
      w.xdrop[l] = take(B * S * (l == 0 ? Aw : (l == 1 ? Aw : Ud) + Aw));
    w.z[l] = take(B * S * 4 * Ud);
    w.c[l] = take(B * S * Ud);
    w.h[l] = take(B * S * Ud);
    w.hprev[l] = take(B * S * Ud);
    if (l + 1 < d.n_layers) w.hdrop[l] = take(B * S * Ud);
    w.dinp[l] = take(2 * B * (2 * (D > Aw ? D : Aw) + Ud));  // 2 slots (bottom_only keeps step t+1's while step t's is written), widest layout
    w.dc[l] = take(B * Ud);
  }
  w.splitk_bytes = (size_t)4 * (d.E + (D > Aw ? D : Aw) + Ud) * 4 * Ud * 4;  // up to 4 K-slices of the largest weight gradient
  w.splitk = take(w.splitk_bytes / 4);
  w.total = off;
  return w;
}

static int gemm(cudaStream_t st, long long M, int N, int K, const float* A, long long sam, long long sak, const float* Bm,
                long long sbk, long long sbn, float* C, long long ldc, const float* bias = nullptr, float beta = 0.f,
                int batch = 1, long long ba = 0, long long bb = 0, long long bc = 0, float* split_ws = nullptr,
                size_t split_bytes = 0) {
  plas_gemm_ex_desc g;
  g.split_ws = split_ws; g.split_ws_bytes = split_bytes;
  g.M = M; g.N = N; g.K = K;
  g.A = A; g.sam = sam; g.sak = sak;
  g.B = Bm; g.sbk = sbk; g.sbn = sbn;
  g.C = C; g.ldc = ldc; g.bias = bias; g.alpha = 1.f; g.beta = beta;
  g.batch = batch; g.batch_a = ba; g.batch_b = bb; g.batch_c = bc;
  return plas_gemm_f32_ex(&g, st);
}

static inline bool att_is_bah(int t) { return t == PLAS_ATT_BAHDANAU || t == PLAS_ATT_BAHDANAU_MONOTONIC; }
static inline bool att_is_mono(int t) { return t == PLAS_ATT_LUONG_MONOTONIC || t == PLAS_ATT_BAHDANAU_MONOTONIC; }

static int dec_train_check(const plas_dec_train_desc* d, void* ws, size_t ws_bytes) {
  PLAS_REQUIRE(d && ws, "dec_train: null argument");
  PLAS_REQUIRE(d->B > 0 && d->S > 0 && d->Tm > 0 && d->E > 0 && d->n_out > 0, "dec_train: bad shape");
  PLAS_REQUIRE(d->n_layers >= 1 && d->n_layers <= 4, "dec_train: n_layers=%d", d->n_layers);
  PLAS_REQUIRE(d->keep_prob > 0.f && d->keep_prob <= 1.f, "dec_train: keep_prob=%f", d->keep_prob);
  PLAS_REQUIRE(d->sample_prob >= 0.f && d->sample_prob <= 1.f, "dec_train: sample_prob=%f", d->sample_prob);
  if (d->sample_prob > 0.f)
    PLAS_REQUIRE(d->x_in_rw != nullptr && (d->E == d->n_out || d->sample_table != nullptr),
                 "dec_train: scheduled sampling needs a writable x_in and one-hot inputs (E == n_out) or the embedding table");
  PLAS_REQUIRE(d->Ud % 16 == 0 && d->D % 4 == 0, "dec_train: Ud=%d must be a multiple of 16, D=%d of 4", d->Ud, d->D);
  PLAS_REQUIRE(d->attention_type >= PLAS_ATT_LUONG && d->attention_type <= PLAS_ATT_CUSTOM, "dec_train: attention_type %d", d->attention_type);
  if (att_is_mono(d->attention_type)) PLAS_REQUIRE(d->score_bias != nullptr, "dec_train: monotonic attention needs attention_score_bias");
  if (att_is_bah(d->attention_type)) PLAS_REQUIRE(d->w_query && d->v_att, "dec_train: bahdanau needs query_layer / attention_v");
  if (d->attention_type == PLAS_ATT_CUSTOM) PLAS_REQUIRE(d->w_query != nullptr, "dec_train: custom attention needs its query_layer");
  PLAS_REQUIRE((size_t)(d->D + 7 * (d->Tm + 4) + 2 * d->Ud) * 4 <= 200 * 1024, "dec_train: D/Tm/Ud too large for one CTA");
  const DecTrainWs w = dec_train_ws(*d);
  PLAS_REQUIRE(ws_bytes >= w.total, "dec_train: workspace %zu < %zu", ws_bytes, w.total);
  return PLAS_OK;
}

// ---- fp32 inference loop ----------------------------------------------------------------------------------------------
struct DecInferWs {
  size_t z[4], c[4], h[4], att, ctx, pq, align, ints, total;
  size_t b_logits, b_logp, b_len, b_ids, b_scratch;  // beam search
};

static DecInferWs dec_infer_ws(const plas_dec_infer_desc& d) {
  DecInferWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off += (bytes + 255) & ~size_t(255);
    return o;
  };
  const size_t B = d.B, Ud = d.Ud, D = d.D, Tm = d.Tm;
  for (int l = 0; l < 4; ++l) {
    w.z[l] = w.c[l] = w.h[l] = 0;
    if (l >= d.n_layers) continue;
    w.z[l] = take(B * 2 * 4 * Ud * 4);
    w.c[l] = take(B * 2 * Ud * 4);
    w.h[l] = take(B * 2 * Ud * 4);
  }
  const size_t A = d.att_layer > 0 ? (size_t)d.att_layer : D;  // width of the attention vector fed back / projected
  w.att = take(B * 2 * A * 4);
  w.ctx = take(B * D * 4);
  w.pq = take(B * Ud * 4);
  w.align = take(2 * B * Tm * 4);  // two slots (parity of t): the monotonic scan reads step t-1's while it writes its own
  w.ints = take((2 * B + 8) * 4);
  w.b_logits = w.b_logp = w.b_len = w.b_ids = w.b_scratch = 0;
  if (d.beam_width > 0) {
    const size_t mw = (Ud > A ? Ud : A) > Tm ? (Ud > A ? Ud : A) : Tm;
    w.b_logits = take(B * (size_t)d.V * 4);
    w.b_logp = take(B * 4);
    w.b_len = take(B * 4);
    w.b_ids = take(B * 4);
    w.b_scratch = take(B * mw * 4);  // out-of-place target of the state gathers
  }
  w.total = off;
  return w;
}

}  // namespace plas

using namespace plas;

extern "C" int plas_relu_f32(float* x, int64_t n, plas_stream_t stream_) {
  PLAS_REQUIRE(x != nullptr && n >= 0, "relu: bad argument");
  if (n) dec_relu_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(x, (size_t)n);
  PLAS_CUDA(cudaGetLastError());
  return PLAS_OK;
}

extern "C" size_t plas_decoder_infer_f32_workspace_bytes(const plas_dec_infer_desc* d) { return dec_infer_ws(*d).total; }

extern "C" int plas_decoder_infer_f32(const plas_dec_infer_desc* d, void* workspace, size_t workspace_bytes, plas_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  PLAS_REQUIRE(d && workspace, "dec_infer: null argument");
  PLAS_REQUIRE(d->B > 0 && d->Tm > 0 && d->V > 0 && d->max_steps >= 0, "dec_infer: bad shape");
  PLAS_REQUIRE(d->n_layers >= 1 && d->n_layers <= 4, "dec_infer: n_layers=%d", d->n_layers);
  PLAS_REQUIRE(d->Ud % 16 == 0 && d->D % 4 == 0, "dec_infer: Ud=%d must be a multiple of 16, D=%d of 4", d->Ud, d->D);
  PLAS_REQUIRE(d->attention_type >= PLAS_ATT_LUONG && d->attention_type <= PLAS_ATT_CUSTOM, "dec_infer: attention_type %d", d->attention_type);
  if (att_is_mono(d->attention_type)) PLAS_REQUIRE(d->score_bias != nullptr, "dec_infer: monotonic attention needs attention_score_bias");
  if (d->attention_type == PLAS_ATT_CUSTOM) PLAS_REQUIRE(d->w_query != nullptr, "dec_infer: custom attention needs its query_layer (and relu'd keys)");
  PLAS_REQUIRE(d->keys && d->values && d->mem_len && d->w_proj && d->b_proj && d->seq_len && d->n_steps, "dec_infer: null tensor");
  PLAS_REQUIRE(d->beam_width > 0 || (d->logits && d->sample_ids), "dec_infer: null logits / sample_ids");
  if (att_is_bah(d->attention_type)) PLAS_REQUIRE(d->w_query && d->v_att, "dec_infer: bahdanau needs query_layer / attention_v");
  if (d->teacher_forced) PLAS_REQUIRE(d->forced_ids != nullptr, "dec_infer: teacher forcing needs forced_ids");
  const DecInferWs w = dec_infer_ws(*d);
  PLAS_REQUIRE(workspace_bytes >= w.total, "dec_infer: workspace %zu < %zu", workspace_bytes, w.total);
  unsigned char* base = (unsigned char*)workspace;
  auto F = [&](size_t off) { return reinterpret_cast<float*>(base + off); };
  const int B = d->B, Tm = d->Tm, D = d->D, Ud = d->Ud, V = d->V, L = d->n_layers, S = d->max_steps;
  InferState is;
  int* ints = reinterpret_cast<int*>(base + w.ints);
  is.cur_ids = ints; is.finished = ints + B; is.done = ints + 2 * B; is.max_iter = ints + 2 * B + 1;
  for (int l = 0; l < L; ++l) {
    PLAS_REQUIRE(d->kernel[l] && d->bias[l], "dec_infer: null weights (layer %d)", l);
    PLAS_CUDA(cudaMemsetAsync(F(w.h[l]), 0, (size_t)B * 2 * Ud * 4, st));
  }
  const int A = d->att_layer > 0 ? d->att_layer : D;
  if (d->att_layer > 0)
    PLAS_REQUIRE(d->w_att_layer && A % 4 == 0, "dec_infer: attention_layer_size needs its kernel and A %% 4 == 0");
  PLAS_CUDA(cudaMemsetAsync(F(w.att), 0, (size_t)B * 2 * A * 4, st));
  dec_infer_init_kernel<<<1, 128, 0, st>>>(is, d->mem_len, B, S, d->teacher_forced, d->decoding_length_factor, d->sos_id, d->seq_len, d->n_steps);
  const int W = d->beam_width;
  BeamState bs = {};
  if (W > 0) {
    PLAS_REQUIRE(W <= 32 && B % W == 0 && !d->teacher_forced && !d->alignment, "dec_infer: beam search needs beam_width <= 32, a tiled batch, no forcing / alignment output");
    PLAS_REQUIRE(d->beam_predicted && d->beam_parent && d->beam_word, "dec_infer: beam search needs beam_predicted / beam_parent / beam_word");
    PLAS_REQUIRE((size_t)W * V * 4 <= 200 * 1024, "dec_infer: beam_width * V too large");
    bs.logp = F(w.b_logp); bs.fin = is.finished; bs.len = reinterpret_cast<int*>(base + w.b_len);
    bs.parent = d->beam_parent; bs.word = d->beam_word;
    dec_beam_init_kernel<<<1, 128, 0, st>>>(is, bs, B, W, d->sos_id, d->seq_len);
    if ((size_t)W * V * 4 > 48 * 1024)
      PLAS_CUDA(cudaFuncSetAttribute(dec_beam_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)W * V * 4)));
  }
  PLAS_CUDA(cudaFuncSetAttribute(dec_cell_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
  PLAS_CUDA(cudaFuncSetAttribute(dec_att_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
  const int dsplit = (D >= 512 && D % 16 == 0) ? 4 : 1;
  size_t att_smem = (size_t)(2 * Ud + ((Tm + 3) & ~3)) * 4;
  const size_t att_stage = (size_t)Tm * Ud * 4 + (size_t)Tm * (D / dsplit) * 4;
  const int att_staged = (Ud % 4 == 0 && (D / dsplit) % 4 == 0 && att_smem + att_stage <= 220 * 1024) ? 1 : 0;
  if (att_staged) att_smem += att_stage;
  const size_t smp_smem = (size_t)((D > Ud ? (D > A ? D : A) : (Ud > A ? Ud : A)) + 4 * V) * 4;  // D >= Ud is not assumed: sized below for the larger of the two
  int rc = PLAS_OK;
  PLAS_REQUIRE(smp_smem <= 200 * 1024, "dec_infer: D/V too large");
  if (smp_smem > 48 * 1024) PLAS_CUDA(cudaFuncSetAttribute(dec_infer_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smp_smem));
  const bool bottom = d->bottom_only != 0;
  const int Dout = (bottom && L > 1) ? Ud : A;  // what the projection reads: the top cell's h (AttentionMultiCell) or the attention
  auto launch_cell = [&](int t, int l) -> int {
    const int slot = t & 1, prev = slot ^ 1;
    CellFwdArgs a;
    a.B = B; a.Ud = Ud; a.skip = is.done;
    // h_{t-1} / c_{t-1} of this layer: the state ring, or the caller's initial state at t = 0 (pass_hidden_state)
    const float* hprev = (t == 0 && d->h_init[l]) ? d->h_init[l] : F(w.h[l]) + (size_t)prev * Ud;
    const long long s_hprev = (t == 0 && d->h_init[l]) ? Ud : 2LL * Ud;
    const float* att_old = F(w.att) + (size_t)prev * A;
    a.in3 = nullptr; a.s3 = 0; a.K3 = 0;
    if (l == 0) {  // [x_t; attention_{t-1}; h_{t-1}]: the x rows are the gathered `pre`
      a.pre = F(w.z[0]) + (size_t)slot * 4 * Ud; a.s_pre = 2LL * 4 * Ud; a.bias = nullptr;
      a.in1 = att_old; a.s1 = 2LL * A; a.K1 = A;
      a.in2 = hprev; a.s2 = s_hprev; a.K2 = Ud;
      a.w = d->kernel[0] + (size_t)V * 4 * Ud;
    } else if (!bottom) {  // MultiRNNCell inside the AttentionWrapper: [h of the layer below; h_{t-1}]
      a.pre = nullptr; a.s_pre = 0; a.bias = d->bias[l];
      a.in1 = F(w.h[l - 1]) + (size_t)slot * Ud; a.s1 = 2LL * Ud; a.K1 = Ud;
      a.in2 = hprev; a.s2 = s_hprev; a.K2 = Ud;
      a.w = d->kernel[l];
    } else {  // AttentionMultiCell: [output of the cell below (cell 0's output is the NEW attention); OLD attention; h_{t-1}]
      a.pre = nullptr; a.s_pre = 0; a.bias = d->bias[l];
      if (l == 1) { a.in1 = F(w.att) + (size_t)slot * A; a.s1 = 2LL * A; a.K1 = A; }
      else { a.in1 = F(w.h[l - 1]) + (size_t)slot * Ud; a.s1 = 2LL * Ud; a.K1 = Ud; }
      a.in2 = att_old; a.s2 = 2LL * A; a.K2 = A;
      a.in3 = hprev; a.s3 = s_hprev; a.K3 = Ud;
      a.w = d->kernel[l];
    }
    a.c_prev = t > 0 ? F(w.c[l]) + (size_t)prev * Ud : d->c_init[l];
    a.s_c = (t == 0) ? Ud : 2LL * Ud;
    a.z_out = F(w.z[l]) + (size_t)slot * 4 * Ud; a.s_z = 2LL * 4 * Ud;
    a.c_out = F(w.c[l]) + (size_t)slot * Ud; a.h_out = F(w.h[l]) + (size_t)slot * Ud; a.s_h = 2LL * Ud;
    a.hprev_next = nullptr; a.hdrop_out = nullptr;
    a.Kdrop = 0; a.dm_stride = a.dm_base = 0; a.dm_seed = 0; a.xdrop = nullptr; a.s_xd = 0;
    a.idx_base = 0; a.seed = 0; a.thresh = 0; a.inv_keep = 1.f; a.step_ptr = nullptr;
    const int per = ((a.K1 + a.K2 + a.K3) / 4 + CF_KS - 1) / CF_KS;
    const size_t smem = ((size_t)DT_ROWS * (4 * per + 4) + (size_t)4 * per * 4 * CF_UG + (size_t)CF_KS * DT_ROWS * 8) * 4;
    PLAS_REQUIRE(smem <= 220 * 1024, "dec_infer: cell input depth %d too large", a.K1 + a.K2 + a.K3);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CF_KS, Ud / CF_UG, (B + DT_ROWS - 1) / DT_ROWS);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CF_KS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PLAS_CUDA(cudaLaunchKernelEx(&cfg, dec_cell_fwd_kernel, a));
    return PLAS_OK;
  };
  auto launch_attention = [&](int t, int query_layer) {
    const int slot = t & 1;
    AttFwdArgs q;
    q.B = B; q.Tm = Tm; q.D = D; q.Ud = Ud; q.type = d->attention_type; q.dsplit = dsplit; q.staged = att_staged; q.skip = is.done;
    q.keys = d->keys; q.values = d->values; q.mem_len = d->mem_len;
    q.query = F(w.h[query_layer]) + (size_t)slot * Ud; q.s_q = 2LL * Ud;
    q.w_query = d->w_query; q.v_att = d->v_att;
    q.pq = F(w.pq); q.s_pq = Ud;
    q.score_bias = d->score_bias; q.p_save = nullptr; q.s_ps = 0;
    q.hard = d->attention_type == PLAS_ATT_BAHDANAU_MONOTONIC ? 1 : 0; q.noise_scale = 0.f; q.noise_seed = 0; q.noise_base = 0;
    if (d->alignment) {
      q.align = d->alignment + (size_t)t * Tm; q.s_al = (long long)S * Tm;
      q.align_prev = t > 0 ? d->alignment + (size_t)(t - 1) * Tm : nullptr; q.s_ap = q.s_al;
    } else {
      q.align = F(w.align) + (size_t)slot * B * Tm; q.s_al = Tm;
      q.align_prev = t > 0 ? F(w.align) + (size_t)(slot ^ 1) * B * Tm : nullptr; q.s_ap = Tm;
    }
    if (d->att_layer > 0) { q.att = F(w.ctx); q.s_att = D; }        // the context goes through the attention layer below
    else { q.att = F(w.att) + (size_t)slot * D; q.s_att = 2LL * D; }
    q.att_next = nullptr; q.next_base = 0; q.seed = 0; q.thresh = 0; q.inv_keep = 1.f; q.step_ptr = nullptr;
    dec_att_fwd_kernel<<<dim3(B, dsplit), 256, att_smem, st>>>(q);
  };
  for (int t = 0; t < S; ++t) {
    const int slot = t & 1;
    dec_infer_embed_kernel<<<(B * 4 * Ud + 255) / 256, 256, 0, st>>>(is, d->forced_ids ? d->forced_ids + t : nullptr, S, d->kernel[0], d->bias[0], B,
                                                                     4 * Ud, V, F(w.z[0]) + (size_t)slot * 4 * Ud, 2LL * 4 * Ud);
    if (bottom) {  // attention right after cell 0 (its query), upper cells afterwards
      if ((rc = launch_cell(t, 0))) return rc;
      launch_attention(t, 0);
      if (d->att_layer > 0) {  // the wrapped cell 0 emits attention = Dense([h0; context]) (A wide)
        float* att_t = F(w.att) + (size_t)slot * A;
        if ((rc = gemm(st, B, A, Ud, F(w.h[0]) + (size_t)slot * Ud, 2LL * Ud, 1, d->w_att_layer, A, 1, att_t, 2LL * A))) return rc;
        if ((rc = gemm(st, B, A, D, F(w.ctx), D, 1, d->w_att_layer + (size_t)Ud * A, A, 1, att_t, 2LL * A, nullptr, 1.f))) return rc;
      }
      for (int l = 1; l < L; ++l)
        if ((rc = launch_cell(t, l))) return rc;
    } else {
      for (int l = 0; l < L; ++l)
        if ((rc = launch_cell(t, l))) return rc;
      launch_attention(t, L - 1);
      if (d->att_layer > 0) {  // attention = Dense([cell output; context]), no bias (AttentionWrapper attention_layer_size)
        float* att_t = F(w.att) + (size_t)slot * A;
        if ((rc = gemm(st, B, A, Ud, F(w.h[L - 1]) + (size_t)slot * Ud, 2LL * Ud, 1, d->w_att_layer, A, 1, att_t, 2LL * A))) return rc;
        if ((rc = gemm(st, B, A, D, F(w.ctx), D, 1, d->w_att_layer + (size_t)Ud * A, A, 1, att_t, 2LL * A, nullptr, 1.f))) return rc;
      }
    }
    const float* out = (bottom && L > 1) ? F(w.h[L - 1]) + (size_t)slot * Ud : F(w.att) + (size_t)slot * A;
    if (W > 0) {  // beam search: logits of every hypothesis, then the W best continuations per utterance and the state gathers
      dec_infer_sample_kernel<<<B, 256, smp_smem, st>>>(is, out, (bottom && L > 1) ? 2LL * Ud : 2LL * A, Dout, V, d->w_proj, d->b_proj,
                                                        F(w.b_logits), V, reinterpret_cast<int*>(base + w.b_ids), 1);
      dec_beam_step_kernel<<<B / W, 256, (size_t)W * V * 4, st>>>(is, bs, F(w.b_logits), B, W, V, t, d->eos_id, d->seq_len);
      const int* parent = d->beam_parent + (size_t)t * B;
      auto regather = [&](float* buf, long long stride, int width) {
        const unsigned blocks = (unsigned)(((size_t)B * width + 255) / 256);
        dec_beam_gather_kernel<<<blocks, 256, 0, st>>>(is, F(w.b_scratch), buf, stride, width, parent, B, W);
        dec_beam_scatter_kernel<<<blocks, 256, 0, st>>>(is, buf, stride, F(w.b_scratch), width, B);
      };
      for (int l = 0; l < L; ++l) {
        regather(F(w.c[l]) + (size_t)slot * Ud, 2LL * Ud, Ud);
        regather(F(w.h[l]) + (size_t)slot * Ud, 2LL * Ud, Ud);
      }
      regather(F(w.att) + (size_t)slot * A, 2LL * A, A);
      if (att_is_mono(d->attention_type)) regather(F(w.align) + (size_t)slot * B * Tm, Tm, Tm);
      dec_beam_finish_kernel<<<1, 32, 0, st>>>(is, B, t, d->n_steps);
      continue;
    }
    dec_infer_sample_kernel<<<B, 256, smp_smem, st>>>(is, out, (bottom && L > 1) ? 2LL * Ud : 2LL * A, Dout, V, d->w_proj, d->b_proj,
                                                      d->logits + (size_t)t * V, (long long)S * V, d->sample_ids + t, S);
    dec_infer_finish_kernel<<<1, 32, 0, st>>>(is, B, t, d->eos_id, d->teacher_forced, d->seq_len, d->n_steps);
  }
  if (W > 0) {
    dec_beam_gather_tree_kernel<<<(B + 127) / 128, 128, 0, st>>>(bs, d->n_steps, B, W, S, d->eos_id, d->beam_predicted);
    if (d->beam_scores) PLAS_CUDA(cudaMemcpyAsync(d->beam_scores, bs.logp, (size_t)B * 4, cudaMemcpyDeviceToDevice, st));
    if (d->beam_lengths) PLAS_CUDA(cudaMemcpyAsync(d->beam_lengths, bs.len, (size_t)B * 4, cudaMemcpyDeviceToDevice, st));
  }
  PLAS_CUDA(cudaGetLastError());
  return PLAS_OK;
}

namespace plas {
// ---------------------------------------------------------------------------------------------------------
// bottom_only training (GNMT-style AttentionMultiCell, las/model.py:20-69, 185-193; README's "true LAS" with pass_hidden_state):
// attention wraps cell 0 only.  Per step: cell 0 ([x_t; attention_{t-1}; h0_{t-1}]) -> attention (query h0_t) -> cells
// l >= 1 ([output below -- the NEW attention for l = 1 --; attention_{t-1}; h_{l,t-1}]) ; logits = h_top W_proj (L > 1).
// Same kernels and saved tensors as the default wiring; only the orchestration differs.
// ---------------------------------------------------------------------------------------------------------
static int launch_cell_fwd(cudaStream_t st, const CellFwdArgs& a, int B, int Ud) {
  const int per = ((a.K1 + a.K2 + a.K3) / 4 + CF_KS - 1) / CF_KS;
  const size_t smem = ((size_t)DT_ROWS * (4 * per + 4) + (size_t)4 * per * 4 * CF_UG + (size_t)CF_KS * DT_ROWS * 8) * 4;
  PLAS_REQUIRE(smem <= 220 * 1024, "decoder cell: input depth %d too large", a.K1 + a.K2 + a.K3);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CF_KS, Ud / CF_UG, (B + DT_ROWS - 1) / DT_ROWS);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CF_KS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  PLAS_CUDA(cudaLaunchKernelEx(&cfg, dec_cell_fwd_kernel, a));
  return PLAS_OK;
}

static int launch_gemv_t(cudaStream_t st, const GemvTArgs& g, int B) {
  const int per = (g.N / 4 + GV_NS - 1) / GV_NS;
  const size_t gsm = ((size_t)DT_ROWS * (4 * per + 4) + (size_t)GV_KG * 4 * per + (size_t)GV_NS * DT_ROWS * 8) * 4;
  PLAS_REQUIRE(gsm <= 220 * 1024, "decoder gemv_t: N = %d too large", g.N);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(GV_NS, (g.K + GV_KG - 1) / GV_KG, (B + DT_ROWS - 1) / DT_ROWS);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = gsm;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = GV_NS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  PLAS_CUDA(cudaLaunchKernelEx(&cfg, dec_gemv_t_kernel, g));
  return PLAS_OK;
}

// launched after the attention (default wiring) / the top cell (bottom_only) of step t when sampling_probability > 0
static int launch_sched_sample(cudaStream_t st, const plas_dec_train_desc* d, const DecTrainWs& w, unsigned char* base, int t, const float* out,
                               long long s_out, int Dout) {
  auto F = [&](size_t off) { return reinterpret_cast<float*>(base + off); };
  const int B = d->B, S = d->S, Ud = d->Ud, E = d->E;
  SchedSampleArgs a;
  a.out = out; a.s_out = s_out; a.Dout = Dout; a.V = d->n_out;
  a.w_proj = d->w_proj; a.b_proj = d->b_proj;
  a.x_next = d->x_in_rw + (size_t)(t + 1) * E; a.s_x = (long long)S * E;
  a.z_next = F(w.z[0]) + (size_t)(t + 1) * 4 * Ud; a.s_z = (long long)S * 4 * Ud; a.N = 4 * Ud;
  a.w0 = d->kernel[0]; a.b0 = d->bias[0];
  a.idx_base = t; a.S = S;
  a.seed_sel = d->sample_seed; a.seed_cat = d->sample_seed + 1; a.p_thresh = (unsigned)(d->sample_prob * 16777216.0f); a.step_ptr = d->drop_step;
  a.xdrop_seed = d->xdrop_seed; a.xdrop_thresh = (unsigned)(d->keep_prob * 16777216.0f); a.xdrop_inv_keep = 1.0f / d->keep_prob;
  a.table = d->sample_table; a.E = E; a.fed_ids = d->sample_fed_ids;
  const size_t smem = (size_t)((Dout > E ? Dout : E) + 4 * d->n_out) * 4;
  PLAS_REQUIRE(smem <= 200 * 1024, "scheduled sampling: projection too large");
  if (smem > 48 * 1024) PLAS_CUDA(cudaFuncSetAttribute(dec_sched_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dec_sched_sample_kernel<<<B, 256, smem, st>>>(a);
  return PLAS_OK;
}

static int dec_train_fwd_bottom(const plas_dec_train_desc* d, void* workspace, cudaStream_t st) {
  const DecTrainWs w = dec_train_ws(*d);
  unsigned char* base = (unsigned char*)workspace;
  auto F = [&](size_t off) { return reinterpret_cast<float*>(base + off); };
  const int B = d->B, S = d->S, Tm = d->Tm, D = d->D, Ud = d->Ud, E = d->E, L = d->n_layers;
  const int AL = d->att_layer, A = AL > 0 ? AL : D;  // attention_layer_size: the wrapped cell 0 emits Dense([h0; context]), A wide
  if (AL > 0) PLAS_REQUIRE(d->w_att_layer && A % 4 == 0, "dec_train: attention_layer_size needs its kernel and A %% 4 == 0");
  const bool custom = d->attention_type == PLAS_ATT_CUSTOM;
  const bool drop = d->keep_prob < 1.f;
  const unsigned thresh = (unsigned)(d->keep_prob * 16777216.0f);
  const float inv_keep = 1.0f / d->keep_prob;

  int rc;
  if ((rc = gemm(st, (long long)B * Tm, Ud, D, d->memory, D, 1, d->w_mem, Ud, 1, F(w.keys), Ud, nullptr, 0.f, 1, 0, 0, 0, F(w.splitk), w.splitk_bytes))) return rc;
  if (custom) dec_relu_kernel<<<(unsigned)(((size_t)B * Tm * Ud + 255) / 256), 256, 0, st>>>(F(w.keys), (size_t)B * Tm * Ud);
  if ((rc = gemm(st, (long long)B * S, 4 * Ud, E, d->x_in, E, 1, d->kernel[0], 4 * Ud, 1, F(w.z[0]), 4 * Ud, d->bias[0]))) return rc;
  PLAS_CUDA(cudaMemset2DAsync(F(w.att_prev), (size_t)S * A * 4, 0, (size_t)A * 4, B, st));
  for (int l = 0; l < L; ++l) {  // h_{-1}: zeros or the listener's final state (pass_hidden_state, las/model.py:259-267)
    if (d->h_init[l]) PLAS_CUDA(cudaMemcpy2DAsync(F(w.hprev[l]), (size_t)S * Ud * 4, d->h_init[l], (size_t)Ud * 4, (size_t)Ud * 4, B, cudaMemcpyDeviceToDevice, st));
    else PLAS_CUDA(cudaMemset2DAsync(F(w.hprev[l]), (size_t)S * Ud * 4, 0, (size_t)Ud * 4, B, st));
  }
  PLAS_CUDA(cudaFuncSetAttribute(dec_cell_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
  PLAS_CUDA(cudaFuncSetAttribute(dec_att_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
  const int dsplit = (D >= 512 && D % 16 == 0) ? 4 : 1;
  size_t att_smem = (size_t)(2 * Ud + ((Tm + 3) & ~3)) * 4;
  const size_t att_stage = (size_t)Tm * Ud * 4 + (size_t)Tm * (D / dsplit) * 4;
  const int att_staged = (Ud % 4 == 0 && (D / dsplit) % 4 == 0 && att_smem + att_stage <= 220 * 1024) ? 1 : 0;
  if (att_staged) att_smem += att_stage;
  const long long sz = (long long)S * 4 * Ud, sh = (long long)S * Ud, sd = (long long)S * D, sa = (long long)S * A;
  for (int t = 0; t < S; ++t) {
    for (int l = 0; l < L; ++l) {
      CellFwdArgs a;
      a.B = B; a.Ud = Ud; a.skip = nullptr;
      a.in3 = nullptr; a.s3 = 0; a.K3 = 0;
      if (l == 0) {
        a.pre = F(w.z[0]) + (size_t)t * 4 * Ud; a.s_pre = sz; a.bias = nullptr;
        a.in1 = F(w.att_prev) + (size_t)t * A; a.s1 = sa; a.K1 = A;
        a.in2 = F(w.hprev[0]) + (size_t)t * Ud; a.s2 = sh; a.K2 = Ud;
        a.w = d->kernel[0] + (size_t)E * 4 * Ud;
      } else {
        a.pre = nullptr; a.s_pre = 0; a.bias = d->bias[l];
        if (l == 1) { a.in1 = F(w.att) + (size_t)t * A; a.s1 = sa; a.K1 = A; }
        else { a.in1 = F(w.h[l - 1]) + (size_t)t * Ud; a.s1 = sh; a.K1 = Ud; }
        a.in2 = F(w.att_prev) + (size_t)t * A; a.s2 = sa; a.K2 = A;
        a.in3 = F(w.hprev[l]) + (size_t)t * Ud; a.s3 = sh; a.K3 = Ud;
        a.w = d->kernel[l];
      }
      if (t > 0) { a.c_prev = F(w.c[l]) + (size_t)(t - 1) * Ud; a.s_c = sh; }
      else { a.c_prev = d->c_init[l]; a.s_c = Ud; }
      a.z_out = F(w.z[l]) + (size_t)t * 4 * Ud; a.s_z = sz;
      a.c_out = F(w.c[l]) + (size_t)t * Ud; a.h_out = F(w.h[l]) + (size_t)t * Ud; a.s_h = sh;
      a.hprev_next = t + 1 < S ? F(w.hprev[l]) + (size_t)(t + 1) * Ud : nullptr;
      a.hdrop_out = nullptr; a.idx_base = 0; a.seed = 0; a.thresh = thresh; a.inv_keep = inv_keep; a.step_ptr = d->drop_step;
      a.Kdrop = 0; a.dm_stride = a.dm_base = 0; a.dm_seed = 0; a.xdrop = nullptr; a.s_xd = 0;
      if (drop) {  // input dropout of cell l: seed drop_seed + l, mask tensor [B][S][Kdrop]
        a.Kdrop = l == 0 ? A : a.K1 + a.K2;
        a.dm_stride = (long long)S * a.Kdrop; a.dm_base = (long long)t * a.Kdrop; a.dm_seed = d->drop_seed + (unsigned)l;
        a.xdrop = F(w.xdrop[l]) + (size_t)t * a.Kdrop; a.s_xd = a.dm_stride;
      }
      if ((rc = launch_cell_fwd(st, a, B, Ud))) return rc;
      if (l == 0) {  // the attention follows cell 0 (its query) and feeds cell 1 of the same step
        AttFwdArgs q;
        q.B = B; q.Tm = Tm; q.D = D; q.Ud = Ud; q.type = d->attention_type; q.dsplit = dsplit; q.staged = att_staged; q.skip = nullptr;
        q.keys = F(w.keys); q.values = d->memory; q.mem_len = d->mem_len;
        q.query = F(w.h[0]) + (size_t)t * Ud; q.s_q = sh;
        q.w_query = d->w_query; q.v_att = d->v_att;
        q.pq = F(w.pq) + (size_t)t * Ud; q.s_pq = sh;
        q.align = F(w.align) + (size_t)t * Tm; q.s_al = (long long)S * Tm;
        q.score_bias = d->score_bias; q.align_prev = t > 0 ? F(w.align) + (size_t)(t - 1) * Tm : nullptr; q.s_ap = q.s_al;
        q.p_save = F(w.psave) + (size_t)t * Tm; q.s_ps = q.s_al;
        q.hard = 0; q.noise_scale = d->attention_type == PLAS_ATT_BAHDANAU_MONOTONIC ? d->sigmoid_noise : 0.f;
        q.noise_seed = d->noise_seed; q.noise_base = (long long)t * Tm;
        if (AL > 0) { q.att = F(w.ctx) + (size_t)t * D; q.s_att = sd; q.att_next = nullptr; }  // the context goes through the attention layer
        else { q.att = F(w.att) + (size_t)t * D; q.s_att = sd; q.att_next = t + 1 < S ? F(w.att_prev) + (size_t)(t + 1) * D : nullptr; }
        q.next_base = 0; q.seed = 0; q.thresh = 0; q.inv_keep = 1.f; q.step_ptr = d->drop_step;  // (the noise seed follows the step)
        dec_att_fwd_kernel<<<dim3(B, dsplit), 256, att_smem, st>>>(q);
        if (AL > 0) {  // attention_t = [h0_t; context_t] W_att (no bias), and the copy every cell reads as OLD attention at step t+1
          float* att_t = F(w.att) + (size_t)t * A;
          if ((rc = gemm(st, B, A, Ud, F(w.h[0]) + (size_t)t * Ud, sh, 1, d->w_att_layer, A, 1, att_t, sa))) return rc;
          if ((rc = gemm(st, B, A, D, F(w.ctx) + (size_t)t * D, sd, 1, d->w_att_layer + (size_t)Ud * A, A, 1, att_t, sa, nullptr, 1.f))) return rc;
          if (t + 1 < S)
            PLAS_CUDA(cudaMemcpy2DAsync(F(w.att_prev) + (size_t)(t + 1) * A, (size_t)sa * 4, att_t, (size_t)sa * 4, (size_t)A * 4, B,
                                        cudaMemcpyDeviceToDevice, st));
        }
      }
    }
    if (d->sample_prob > 0.f && t + 1 < S) {
      if (L > 1) rc = launch_sched_sample(st, d, w, base, t, F(w.h[L - 1]) + (size_t)t * Ud, sh, Ud);
      else rc = launch_sched_sample(st, d, w, base, t, F(w.att) + (size_t)t * A, sa, A);
      if (rc) return rc;
    }
  }
  PLAS_CUDA(cudaGetLastError());
  if (L > 1) return gemm(st, (long long)B * S, d->n_out, Ud, F(w.h[L - 1]), Ud, 1, d->w_proj, d->n_out, 1, d->logits, d->n_out, d->b_proj);
  return gemm(st, (long long)B * S, d->n_out, A, F(w.att), A, 1, d->w_proj, d->n_out, 1, d->logits, d->n_out, d->b_proj, 0.f, 1, 0, 0, 0,
              F(w.splitk), w.splitk_bytes);
}

static int dec_train_bwd_bottom(const plas_dec_train_desc* d, void* workspace, cudaStream_t st) {
  const DecTrainWs w = dec_train_ws(*d);
  unsigned char* base = (unsigned char*)workspace;
  auto F = [&](size_t off) { return reinterpret_cast<float*>(base + off); };
  const int B = d->B, S = d->S, Tm = d->Tm, D = d->D, Ud = d->Ud, E = d->E, L = d->n_layers, NO = d->n_out;
  const bool bah = att_is_bah(d->attention_type);
  const bool custom = d->attention_type == PLAS_ATT_CUSTOM;
  const long long BS = (long long)B * S;
  const long long sz = (long long)S * 4 * Ud, sh = (long long)S * Ud, sd = (long long)S * D;
  float* dctx = F(w.dctx);
  const int AL = d->att_layer, A = AL > 0 ? AL : D;
  float* datt = AL > 0 ? F(w.datt) : dctx;     // gradient wrt the attention vectors [B][S][A]; without the layer it IS dctx
  if (AL > 0) PLAS_REQUIRE(d->w_att_layer && d->dw_att_layer, "dec_train_bwd: attention_layer_size needs w_att_layer / dw_att_layer");
  const long long sa = (long long)S * A;
  float* sk = F(w.splitk);
  const size_t skb = w.splitk_bytes;
  const bool drop = d->keep_prob < 1.f;
  const unsigned thresh = (unsigned)(d->keep_prob * 16777216.0f);
  const float inv_keep = 1.0f / d->keep_prob;
  int rc;
  // projection: reads the top cell's h (L > 1) or the attention (L == 1)
  if (L > 1) {
    if ((rc = gemm(st, BS, Ud, NO, d->dlogits, NO, 1, d->w_proj, 1, NO, F(w.gh), Ud))) return rc;
    if ((rc = gemm(st, Ud, NO, (int)BS, F(w.h[L - 1]), 1, Ud, d->dlogits, NO, 1, d->dw_proj, NO))) return rc;
    PLAS_CUDA(cudaMemsetAsync(datt, 0, (size_t)BS * A * 4, st));
  } else {
    if ((rc = gemm(st, BS, A, NO, d->dlogits, NO, 1, d->w_proj, 1, NO, datt, A))) return rc;
    if ((rc = gemm(st, A, NO, (int)BS, F(w.att), 1, A, d->dlogits, NO, 1, d->dw_proj, NO))) return rc;
  }
  if ((rc = plas_colsum_f32(d->dlogits, BS, NO, NO, d->db_proj, 0, st))) return rc;
  if (bah) {
    PLAS_REQUIRE(d->dw_query && d->dv_att, "dec_train_bwd: bahdanau needs dw_query / dv_att");
    PLAS_CUDA(cudaMemsetAsync(F(w.dkeys), 0, (size_t)B * Tm * Ud * 4, st));
    PLAS_CUDA(cudaMemsetAsync(F(w.dv_acc), 0, (size_t)B * Ud * 4, st));
  }
  const bool mono = att_is_mono(d->attention_type);
  if (mono) {
    PLAS_REQUIRE(d->dscore_bias != nullptr, "dec_train_bwd: luong_monotonic needs dscore_bias");
    PLAS_CUDA(cudaMemsetAsync(F(w.dbias), 0, (size_t)B * 4, st));
  }
  const int ns = (D % 16 == 0 && Ud % 16 == 0 && D >= 256) ? 4 : 1;
  const int Tp = (Tm + 3) & ~3;
  size_t attb_smem = (size_t)(D / ns + (ns + 3) * Tp + Ud) * 4;
  const size_t attb_stage = (size_t)Tm * (D / ns + Ud / ns) * 4;
  const int attb_staged = ((D / ns) % 4 == 0 && (Ud / ns) % 4 == 0 && attb_smem + attb_stage <= 220 * 1024) ? 1 : 0;
  if (attb_staged) attb_smem += attb_stage;
  PLAS_CUDA(cudaFuncSetAttribute(dec_att_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
  PLAS_CUDA(cudaFuncSetAttribute(dec_gemv_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
  // dinp_l = dz_l W_l^T lives in two slots (parity of t): step t reads step t+1's while it writes its own
  auto Kin = [&](int l) { return l == 0 ? A : ((l == 1 ? A : Ud) + A); };        // input columns before the h part
  auto dinp = [&](int l, int t) { return F(w.dinp[l]) + (size_t)(t & 1) * B * (2 * (D > A ? D : A) + Ud); };
  for (int t = S - 1; t >= 0; --t) {
    const bool last = t == S - 1;
    auto cell_and_gemv = [&](int l) -> int {
      CellBwdArgs c;
      c.B = B; c.Ud = Ud;
      c.z = F(w.z[l]) + (size_t)t * 4 * Ud; c.s_z = sz;
      c.c_t = F(w.c[l]) + (size_t)t * Ud; c.s_c = sh;
      if (t > 0) { c.c_prev = F(w.c[l]) + (size_t)(t - 1) * Ud; c.s_cp = sh; }
      else { c.c_prev = d->c_init[l]; c.s_cp = Ud; }
      c.dc_carry = F(w.dc[l]); c.first = last ? 1 : 0;
      if (l == 0) { c.dq = F(w.dq); c.s_dq = Ud; }                                          // gradient wrt the query h0_t
      else if (l == L - 1) { c.dq = F(w.gh) + (size_t)t * Ud; c.s_dq = sh; }                // the projection reads the top cell
      else { c.dq = nullptr; c.s_dq = 0; }
      c.dh_next = last ? nullptr : dinp(l, t + 1) + Kin(l); c.s_dn = Kin(l) + Ud;
      // the layer above reads this layer's h as its first input segment -- except above cell 0, whose output is the attention
      c.dh_above = (l >= 1 && l < L - 1) ? dinp(l + 1, t) : nullptr; c.s_da = Kin(l + 1 < L ? l + 1 : l) + Ud;
      c.dh_extra = (l == 0 && AL > 0) ? F(w.dqx) : nullptr; c.s_dx = Ud;  // the attention layer also reads h0
      c.idx_base = 0; c.seed = 0; c.thresh = 0; c.inv_keep = 1.f; c.step_ptr = nullptr;
      dec_cell_bwd_kernel<<<(B * Ud + 255) / 256, 256, 0, st>>>(c);
      GemvTArgs g;
      g.B = B; g.N = 4 * Ud; g.K = Kin(l) + Ud;
      g.dz = c.z; g.s_z = sz;
      g.w = d->kernel[l] + (size_t)(l == 0 ? E : 0) * 4 * Ud;
      g.dinp = dinp(l, t); g.s_o = g.K;
      if ((rc = launch_gemv_t(st, g, B))) return rc;
      if (drop)  // back through the cell's input dropout: every reader of dinp's input columns then sees the masked gradient
        dec_add_rows_kernel<<<(B * Kin(l) + 255) / 256, 256, 0, st>>>(dinp(l, t), g.K, nullptr, 0, dinp(l, t), g.K, B, Kin(l), (long long)S * Kin(l),
                                                                      (long long)t * Kin(l), d->drop_seed + (unsigned)l, thresh, inv_keep, d->drop_step);
      return PLAS_OK;
    };
    for (int l = L - 1; l >= 1; --l)
      if ((rc = cell_and_gemv(l))) return rc;
    AttBwdArgs q;
    q.B = B; q.Tm = Tm; q.D = D; q.Ud = Ud; q.type = d->attention_type; q.nsplit = ns; q.staged = attb_staged;
    q.keys = F(w.keys); q.values = d->memory; q.mem_len = d->mem_len;
    q.align = F(w.align) + (size_t)t * Tm; q.s_al = (long long)S * Tm;
    q.dctx = dctx + (size_t)t * D; q.s_dc = sd;
    q.datt_next = nullptr; q.s_dn = 0;
    for (int e = 0; e < 4; ++e) { q.extra[e] = nullptr; q.s_extra[e] = 0; }
    if (AL > 0) {
      // gradient wrt attention_t (A wide): the projection (L == 1, already in datt) + cell 0's input at step t+1 + cell 1's first
      // input at this step + every upper cell's OLD attention at step t+1; then back through attention = [h0; context] W_att
      float* da = datt + (size_t)t * A;
      auto add = [&](const float* src, long long stride) {
        dec_add_rows_kernel<<<(B * A + 255) / 256, 256, 0, st>>>(da, sa, da, sa, src, stride, B, A);
      };
      if (!last) add(dinp(0, t + 1), Kin(0) + Ud);
      if (L > 1) add(dinp(1, t), Kin(1) + Ud);
      if (!last)
        for (int l = 1; l < L; ++l) add(dinp(l, t + 1) + (l == 1 ? A : Ud), Kin(l) + Ud);
      if ((rc = gemm(st, B, Ud, A, da, sa, 1, d->w_att_layer, 1, A, F(w.dqx), Ud))) return rc;
      if ((rc = gemm(st, B, D, A, da, sa, 1, d->w_att_layer + (size_t)Ud * A, 1, A, dctx + (size_t)t * D, sd))) return rc;
    } else {
      q.datt_next = last ? nullptr : dinp(0, t + 1); q.s_dn = D + Ud;      // attention_t was cell 0's input at step t+1
      int ne = 0;
      if (L > 1) { q.extra[ne] = dinp(1, t); q.s_extra[ne] = Kin(1) + Ud; ++ne; }   // ... cell 1's first input at this step
      if (!last)
        for (int l = 1; l < L && ne < 4; ++l) {                                      // ... and every upper cell's OLD attention at step t+1
          q.extra[ne] = dinp(l, t + 1) + (l == 1 ? D : Ud); q.s_extra[ne] = Kin(l) + Ud; ++ne;
        }
    }
    q.dscore = F(w.dscore) + (size_t)t * Tm; q.s_ds = (long long)S * Tm;
    q.score_p = F(w.psave) + (size_t)t * Tm; q.s_sp = q.s_al;
    q.align_prev = t > 0 ? F(w.align) + (size_t)(t - 1) * Tm : nullptr; q.s_ap = q.s_al;
    q.dalign_in = last ? nullptr : F(w.dalign) + (size_t)((t + 1) & 1) * B * Tm;
    q.dalign_out = F(w.dalign) + (size_t)(t & 1) * B * Tm; q.dbias_acc = F(w.dbias);
    q.dq = F(w.dq); q.s_dq = Ud;
    q.pq = F(w.pq) + (size_t)t * Ud; q.s_pq = sh; q.w_query = d->w_query; q.v_att = d->v_att;
    q.dpq = F(w.dpq) + (size_t)t * Ud; q.s_dpq = sh; q.dkeys = F(w.dkeys); q.dv_acc = F(w.dv_acc);
    q.next_base = 0; q.seed = 0; q.thresh = 0; q.inv_keep = 1.f; q.step_ptr = nullptr;
    {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(B, ns);
      cfg.blockDim = dim3(256);
      cfg.dynamicSmemBytes = attb_smem;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 1;
      attr[0].val.clusterDim.y = (unsigned)ns;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      PLAS_CUDA(cudaLaunchKernelEx(&cfg, dec_att_bwd_kernel, q));
    }
    if ((rc = cell_and_gemv(0))) return rc;
  }
  PLAS_CUDA(cudaGetLastError());
  // gradients wrt the initial states (flow into the listener's final states with pass_hidden_state)
  for (int l = 0; l < L; ++l) {
    if (d->dh_init[l])
      PLAS_CUDA(cudaMemcpy2DAsync(d->dh_init[l], (size_t)Ud * 4, dinp(l, 0) + Kin(l), (size_t)(Kin(l) + Ud) * 4, (size_t)Ud * 4, B,
                                  cudaMemcpyDeviceToDevice, st));
    if (d->dc_init[l]) PLAS_CUDA(cudaMemcpyAsync(d->dc_init[l], F(w.dc[l]), (size_t)B * Ud * 4, cudaMemcpyDeviceToDevice, st));
  }
  // weight gradients over the B*S saved rows (z now holds dz)
  for (int l = 0; l < L; ++l) {
    PLAS_REQUIRE(d->dkernel[l] && d->dbias[l], "dec_train_bwd: null gradient tensor (layer %d)", l);
    const float* dz = F(w.z[l]);
    float* dk = d->dkernel[l];
    auto wg = [&](const float* a, int rows, size_t row0) {
      return gemm(st, rows, 4 * Ud, (int)BS, a, 1, rows, dz, 4 * Ud, 1, dk + row0 * 4 * Ud, 4 * Ud, nullptr, 0.f, 1, 0, 0, 0, sk, skb);
    };
    if (l == 0) {
      if ((rc = wg(d->x_in, E, 0))) return rc;
      if (d->dx_in && (rc = gemm(st, BS, E, 4 * Ud, dz, 4 * Ud, 1, d->kernel[0], 1, 4 * Ud, d->dx_in, E))) return rc;  // dX = dZ_0 W_0[0:E]^T
      if ((rc = wg(drop ? F(w.xdrop[0]) : F(w.att_prev), A, E))) return rc;
      if ((rc = wg(F(w.hprev[0]), Ud, (size_t)E + A))) return rc;
    } else {
      const int k1 = l == 1 ? A : Ud;
      if (drop) {  // the dropped-out [output below; old attention] rows the forward pass kept
        if ((rc = wg(F(w.xdrop[l]), k1 + A, 0))) return rc;
      } else {
        if ((rc = wg(l == 1 ? F(w.att) : F(w.h[l - 1]), k1, 0))) return rc;
        if ((rc = wg(F(w.att_prev), A, k1))) return rc;
      }
      if ((rc = wg(F(w.hprev[l]), Ud, (size_t)k1 + A))) return rc;
    }
    if ((rc = plas_colsum_f32(dz, BS, 4 * Ud, 4 * Ud, d->dbias[l], 0, st))) return rc;
  }
  const float* h0 = F(w.h[0]);  // the query is cell 0's output
  if (AL > 0) {  // dW_att = [H0; Ctx]^T dAtt
    if ((rc = gemm(st, Ud, A, (int)BS, h0, 1, Ud, datt, A, 1, d->dw_att_layer, A, nullptr, 0.f, 1, 0, 0, 0, sk, skb))) return rc;
    if ((rc = gemm(st, D, A, (int)BS, F(w.ctx), 1, D, datt, A, 1, d->dw_att_layer + (size_t)Ud * A, A, nullptr, 0.f, 1, 0, 0, 0, sk, skb))) return rc;
  }
  if (bah) {
    if ((rc = gemm(st, Ud, Ud, (int)BS, h0, 1, Ud, F(w.dpq), Ud, 1, d->dw_query, Ud))) return rc;
    if ((rc = plas_colsum_f32(F(w.dv_acc), B, Ud, Ud, d->dv_att, 0, st))) return rc;
  } else {  // dkeys[b] = dScore[b]^T Q[b], Q = the query the score read (custom: relu(query_layer(h)), saved in pq)
    if ((rc = gemm(st, Tm, Ud, S, F(w.dscore), 1, Tm, custom ? F(w.pq) : h0, Ud, 1, F(w.dkeys), Ud, nullptr, 0.f, B, (long long)S * Tm,
                   (long long)S * Ud, (long long)Tm * Ud)))
      return rc;
    if (custom) {  // keys = relu(memory_layer(values)); dW_query = H^T dPQ
      dec_relu_grad_kernel<<<(unsigned)(((size_t)B * Tm * Ud + 255) / 256), 256, 0, st>>>(F(w.dkeys), F(w.keys), (size_t)B * Tm * Ud);
      PLAS_REQUIRE(d->dw_query != nullptr, "dec_train_bwd: custom attention needs dw_query");
      if ((rc = gemm(st, Ud, Ud, (int)BS, h0, 1, Ud, F(w.dpq), Ud, 1, d->dw_query, Ud))) return rc;
    }
  }
  if ((rc = gemm(st, Tm, D, S, F(w.align), 1, Tm, dctx, D, 1, d->dmemory, D, nullptr, d->dmemory_accumulate ? 1.f : 0.f, B, (long long)S * Tm,
                 (long long)S * D, (long long)Tm * D)))
    return rc;
  if (mono && (rc = plas_colsum_f32(F(w.dbias), B, 1, 1, d->dscore_bias, 0, st))) return rc;
  if ((rc = gemm(st, (long long)B * Tm, D, Ud, F(w.dkeys), Ud, 1, d->w_mem, 1, Ud, d->dmemory, D, nullptr, 1.f))) return rc;
  return gemm(st, D, Ud, B * Tm, d->memory, 1, D, F(w.dkeys), Ud, 1, d->dw_mem, Ud, nullptr, 0.f, 1, 0, 0, 0, sk, skb);
}

}  // namespace plas

using namespace plas;

extern "C" size_t plas_dec_train_workspace_bytes(const plas_dec_train_desc* d) { return dec_train_ws(*d).total; }

extern "C" int plas_decoder_train_fwd(const plas_dec_train_desc* d, void* workspace, size_t workspace_bytes,
                                      plas_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  int rc = dec_train_check(d, workspace, workspace_bytes);
  if (rc) return rc;
  PLAS_REQUIRE(d->memory && d->mem_len && d->x_in && d->logits && d->w_mem && d->w_proj && d->b_proj, "dec_train_fwd: null tensor");
  if (d->bottom_only) return dec_train_fwd_bottom(d, workspace, st);
  const DecTrainWs w = dec_train_ws(*d);
  unsigned char* base = (unsigned char*)workspace;
  auto F = [&](size_t off) { return reinterpret_cast<float*>(base + off); };
  const int B = d->B, S = d->S, Tm = d->Tm, D = d->D, Ud = d->Ud, E = d->E, L = d->n_layers;
  const int A = d->att_layer > 0 ? d->att_layer : D;  // width of the attention vector (attention_layer_size or the context depth)
  if (d->att_layer > 0)
    PLAS_REQUIRE(d->w_att_layer && A % 4 == 0, "dec_train: attention_layer_size needs its kernel and A %% 4 == 0");
  const bool bah = att_is_bah(d->attention_type);
  const bool custom = d->attention_type == PLAS_ATT_CUSTOM;
  if (bah) PLAS_REQUIRE(d->w_query && d->v_att, "dec_train_fwd: bahdanau needs query_layer / attention_v");
  // keys = memory_layer(values); the memory is already zero past each length (the listener guarantees it)
  // (few output tiles, long contraction: deterministic split-K through the workspace's scratch, like the weight gradients)
  rc = gemm(st, (long long)B * Tm, Ud, D, d->memory, D, 1, d->w_mem, Ud, 1, F(w.keys), Ud, nullptr, 0.f, 1, 0, 0, 0, F(w.splitk), w.splitk_bytes);
  if (rc) return rc;
  if (custom) dec_relu_kernel<<<(unsigned)(((size_t)B * Tm * Ud + 255) / 256), 256, 0, st>>>(F(w.keys), (size_t)B * Tm * Ud);
  // Z_0 = x_in W_0[0:E] + b_0 for every step at once
  rc = gemm(st, (long long)B * S, 4 * Ud, E, d->x_in, E, 1, d->kernel[0], 4 * Ud, 1, F(w.z[0]), 4 * Ud, d->bias[0]);
  if (rc) return rc;
  PLAS_CUDA(cudaMemset2DAsync(F(w.att_prev), (size_t)S * A * 4, 0, (size_t)A * 4, B, st));
  for (int l = 0; l < L; ++l) PLAS_CUDA(cudaMemset2DAsync(F(w.hprev[l]), (size_t)S * Ud * 4, 0, (size_t)Ud * 4, B, st));
  const bool drop = d->keep_prob < 1.f;
  const unsigned thresh = (unsigned)(d->keep_prob * 16777216.0f);
  const float inv_keep = 1.0f / d->keep_prob;
  PLAS_CUDA(cudaFuncSetAttribute(dec_cell_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
  const int dsplit = (D >= 512 && D % 16 == 0) ? 4 : 1;
  size_t att_smem = (size_t)(2 * Ud + ((Tm + 3) & ~3)) * 4;
  const size_t att_stage = (size_t)Tm * Ud * 4 + (size_t)Tm * (D / dsplit) * 4;
  const int att_staged = (Ud % 4 == 0 && (D / dsplit) % 4 == 0 && att_smem + att_stage <= 220 * 1024) ? 1 : 0;
  if (att_staged) att_smem += att_stage;
  PLAS_CUDA(cudaFuncSetAttribute(dec_att_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
  for (int t = 0; t < S; ++t) {
    for (int l = 0; l < L; ++l) {
      CellFwdArgs a;
      a.B = B; a.Ud = Ud;
      const long long sz = (long long)S * 4 * Ud, sh = (long long)S * Ud;
      if (l == 0) {
        a.pre = F(w.z[0]) + (size_t)t * 4 * Ud; a.s_pre = sz; a.bias = nullptr;
        a.in1 = F(w.att_prev) + (size_t)t * A; a.s1 = (long long)S * A; a.K1 = A;
        a.w = d->kernel[0] + (size_t)E * 4 * Ud;  // rows: [x (E, hoisted); attention (A); h (Ud)]
      } else {
        a.pre = nullptr; a.s_pre = 0; a.bias = d->bias[l];
        a.in1 = F(drop ? w.hdrop[l - 1] : w.h[l - 1]) + (size_t)t * Ud; a.s1 = sh; a.K1 = Ud;
        a.w = d->kernel[l];
      }
      a.in2 = F(w.hprev[l]) + (size_t)t * Ud; a.s2 = sh; a.K2 = Ud;
      a.in3 = nullptr; a.s3 = 0; a.K3 = 0;
      a.c_prev = t > 0 ? F(w.c[l]) + (size_t)(t - 1) * Ud : nullptr; a.s_c = sh;
      a.z_out = F(w.z[l]) + (size_t)t * 4 * Ud; a.s_z = sz;
      a.c_out = F(w.c[l]) + (size_t)t * Ud; a.h_out = F(w.h[l]) + (size_t)t * Ud; a.s_h = sh;
      a.skip = nullptr;
      a.hprev_next = t + 1 < S ? F(w.hprev[l]) + (size_t)(t + 1) * Ud : nullptr;
      a.hdrop_out = (drop && l + 1 < L) ? F(w.hdrop[l]) + (size_t)t * Ud : nullptr;
      a.Kdrop = 0; a.dm_stride = a.dm_base = 0; a.dm_seed = 0; a.xdrop = nullptr; a.s_xd = 0;
      a.idx_base = (long long)t * Ud; a.seed = d->drop_seed + 1 + l; a.thresh = thresh; a.inv_keep = inv_keep; a.step_ptr = d->drop_step;
      {
        const int per = ((a.K1 + a.K2) / 4 + CF_KS - 1) / CF_KS;
        const size_t smem = ((size_t)DT_ROWS * (4 * per + 4) + (size_t)4 * per * 4 * CF_UG + (size_t)CF_KS * DT_ROWS * 8) * 4;
        PLAS_REQUIRE(smem <= 220 * 1024, "dec_train_fwd: cell input depth %d too large", a.K1 + a.K2);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(CF_KS, Ud / CF_UG, (B + DT_ROWS - 1) / DT_ROWS);
        cfg.blockDim = dim3(256);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CF_KS;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        PLAS_CUDA(cudaLaunchKernelEx(&cfg, dec_cell_fwd_kernel, a));
      }
    }
    AttFwdArgs q;
    q.B = B; q.Tm = Tm; q.D = D; q.Ud = Ud; q.type = d->attention_type; q.dsplit = dsplit; q.staged = att_staged;
    q.keys = F(w.keys); q.values = d->memory; q.mem_len = d->mem_len;
    q.query = F(w.h[L - 1]) + (size_t)t * Ud; q.s_q = (long long)S * Ud;
    q.w_query = d->w_query; q.v_att = d->v_att;
    q.pq = F(w.pq) + (size_t)t * Ud; q.s_pq = (long long)S * Ud;
    q.align = F(w.align) + (size_t)t * Tm; q.s_al = (long long)S * Tm;
    q.score_bias = d->score_bias; q.align_prev = t > 0 ? F(w.align) + (size_t)(t - 1) * Tm : nullptr; q.s_ap = q.s_al;
    q.p_save = F(w.psave) + (size_t)t * Tm; q.s_ps = q.s_al;
    q.hard = 0; q.noise_scale = d->attention_type == PLAS_ATT_BAHDANAU_MONOTONIC ? d->sigmoid_noise : 0.f;
    q.noise_seed = d->noise_seed; q.noise_base = (long long)t * Tm;
    if (d->att_layer > 0) { q.att = F(w.ctx) + (size_t)t * D; q.s_att = (long long)S * D; }
    else { q.att = F(w.att) + (size_t)t * D; q.s_att = (long long)S * D; }
    q.skip = nullptr;
    q.att_next = (t + 1 < S && d->att_layer == 0) ? F(w.att_prev) + (size_t)(t + 1) * D : nullptr;
    q.next_base = (long long)(t + 1) * D; q.seed = d->drop_seed; q.thresh = thresh; q.inv_keep = inv_keep; q.step_ptr = d->drop_step;
    dec_att_fwd_kernel<<<dim3(B, dsplit), 256, att_smem, st>>>(q);
    if (d->att_layer > 0) {  // attention_t = [h_top_t; context_t] W_att (no bias), then the copy that step t+1 reads
      float* att_t = F(w.att) + (size_t)t * A;
      if ((rc = gemm(st, B, A, Ud, F(w.h[L - 1]) + (size_t)t * Ud, (long long)S * Ud, 1, d->w_att_layer, A, 1, att_t, (long long)S * A))) return rc;
      if ((rc = gemm(st, B, A, D, F(w.ctx) + (size_t)t * D, (long long)S * D, 1, d->w_att_layer + (size_t)Ud * A, A, 1, att_t, (long long)S * A,
                     nullptr, 1.f)))
        return rc;
      if (t + 1 < S)  // the (dropped-out) copy that cell 0 reads at step t+1: mask element [b][t+1][a] of the [B][S][A] tensor
        dec_add_rows_kernel<<<(B * A + 255) / 256, 256, 0, st>>>(F(w.att_prev) + (size_t)(t + 1) * A, (long long)S * A, nullptr, 0, att_t,
                                                                 (long long)S * A, B, A, (long long)S * A, (long long)(t + 1) * A, d->drop_seed,
                                                                 thresh, inv_keep, d->drop_step);
    }
    if (d->sample_prob > 0.f && t + 1 < S)
      if ((rc = launch_sched_sample(st, d, w, base, t, F(w.att) + (size_t)t * A, (long long)S * A, A))) return rc;
  }
  PLAS_CUDA(cudaGetLastError());
  if (d->att_out) PLAS_CUDA(cudaMemcpyAsync(d->att_out, F(w.att), (size_t)B * S * A * 4, cudaMemcpyDeviceToDevice, st));
  // logits = DenseBinfDecoder(attention)
  return gemm(st, (long long)B * S, d->n_out, A, F(w.att), A, 1, d->w_proj, d->n_out, 1, d->logits, d->n_out, d->b_proj, 0.f, 1, 0, 0, 0,
              F(w.splitk), w.splitk_bytes);
}

extern "C" int plas_decoder_train_bwd(const plas_dec_train_desc* d, void* workspace, size_t workspace_bytes,
                                      plas_stream_t stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  int rc = dec_train_check(d, workspace, workspace_bytes);
  if (rc) return rc;
  PLAS_REQUIRE(d->dlogits && d->dmemory && d->dw_mem, "dec_train_bwd: null tensor");
  PLAS_REQUIRE(d->bottom_only == 0 || (d->dw_proj && d->db_proj && !d->datt_extra), "dec_train_bwd: bottom_only needs dw_proj / db_proj, no datt_extra");
  if (d->bottom_only) return dec_train_bwd_bottom(d, workspace, st);
  const DecTrainWs w = dec_train_ws(*d);
  unsigned char* base = (unsigned char*)workspace;
  auto F = [&](size_t off) { return reinterpret_cast<float*>(base + off); };
  const int B = d->B, S = d->S, Tm = d->Tm, D = d->D, Ud = d->Ud, E = d->E, L = d->n_layers, NO = d->n_out;
  const bool bah = att_is_bah(d->attention_type);
  const bool custom = d->attention_type == PLAS_ATT_CUSTOM;
  const long long BS = (long long)B * S;
  float* dctx = F(w.dctx);  // [B][S][D]
  const int AL = d->att_layer;                 // attention_layer_size (0 = none)
  const int A = AL > 0 ? AL : D;               // width of the attention vector
  float* datt = AL > 0 ? F(w.datt) : dctx;     // gradient wrt the attention vector ([B][S][A]); without the layer it IS dctx
  if (AL > 0) PLAS_REQUIRE(d->w_att_layer && d->dw_att_layer, "dec_train_bwd: attention_layer_size needs w_att_layer / dw_att_layer");
  // projection layer: dAtt = dlogits W_proj^T, dW_proj = Att^T dlogits, db_proj = colsum(dlogits)
  // (+ datt_extra: a gradient that reaches the attention vectors directly, e.g. the --binf_projection regulariser)
  if (d->datt_extra) PLAS_CUDA(cudaMemcpyAsync(datt, d->datt_extra, (size_t)BS * A * 4, cudaMemcpyDeviceToDevice, st));
  if ((rc = gemm(st, BS, A, NO, d->dlogits, NO, 1, d->w_proj, 1, NO, datt, A, nullptr, d->datt_extra ? 1.f : 0.f))) return rc;
  // dw_proj / db_proj NULL: the projection is a constant (transform_binf_to_phones under --binf_projection)
  if (d->dw_proj && (rc = gemm(st, A, NO, (int)BS, F(w.att), 1, A, d->dlogits, NO, 1, d->dw_proj, NO, nullptr, 0.f, 1, 0, 0, 0, F(w.splitk), w.splitk_bytes))) return rc;
  if (d->db_proj && (rc = plas_colsum_f32(d->dlogits, BS, NO, NO, d->db_proj, 0, st))) return rc;
  if (bah) {
    PLAS_REQUIRE(d->dw_query && d->dv_att, "dec_train_bwd: bahdanau needs dw_query / dv_att");
    PLAS_CUDA(cudaMemsetAsync(F(w.dkeys), 0, (size_t)B * Tm * Ud * 4, st));
    PLAS_CUDA(cudaMemsetAsync(F(w.dv_acc), 0, (size_t)B * Ud * 4, st));
  }
  const bool mono = att_is_mono(d->attention_type);
  if (mono) {
    PLAS_REQUIRE(d->dscore_bias != nullptr, "dec_train_bwd: luong_monotonic needs dscore_bias");
    PLAS_CUDA(cudaMemsetAsync(F(w.dbias), 0, (size_t)B * 4, st));
  }
  const int ns = (D % 16 == 0 && Ud % 16 == 0 && D >= 256) ? 4 : 1;
  const int Tp = (Tm + 3) & ~3;
  size_t attb_smem = (size_t)(D / ns + (ns + 3) * Tp + Ud) * 4;
  const size_t attb_stage = (size_t)Tm * (D / ns + Ud / ns) * 4;
  const int attb_staged = ((D / ns) % 4 == 0 && (Ud / ns) % 4 == 0 && attb_smem + attb_stage <= 220 * 1024) ? 1 : 0;
  if (attb_staged) attb_smem += attb_stage;
  PLAS_CUDA(cudaFuncSetAttribute(dec_att_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
  PLAS_CUDA(cudaFuncSetAttribute(dec_gemv_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  float* sk = F(w.splitk);
  const size_t skb = w.splitk_bytes;
  const long long sz = (long long)S * 4 * Ud, sh = (long long)S * Ud;
  const bool drop = d->keep_prob < 1.f;
  const unsigned thresh = (unsigned)(d->keep_prob * 16777216.0f);
  const float inv_keep = 1.0f / d->keep_prob;
  for (int t = S - 1; t >= 0; --t) {
    const bool last = t == S - 1;
    if (AL > 0) {  // datt_t = dlogits_t W_proj^T + (step t+1's cell-0 input gradient); back through attention = [h_top; ctx] W_att
      float* da = datt + (size_t)t * A;
      if (!last)
        dec_add_rows_kernel<<<(B * A + 255) / 256, 256, 0, st>>>(da, (long long)S * A, da, (long long)S * A, F(w.dinp[0]), A + Ud, B, A,
                                                                 (long long)S * A, (long long)(t + 1) * A, d->drop_seed, thresh, inv_keep, d->drop_step);
      if ((rc = gemm(st, B, Ud, A, da, (long long)S * A, 1, d->w_att_layer, 1, A, F(w.dqx), Ud))) return rc;
      if ((rc = gemm(st, B, D, A, da, (long long)S * A, 1, d->w_att_layer + (size_t)Ud * A, 1, A, dctx + (size_t)t * D, (long long)S * D))) return rc;
    }
    AttBwdArgs q;
    q.B = B; q.Tm = Tm; q.D = D; q.Ud = Ud; q.type = d->attention_type; q.nsplit = ns; q.staged = attb_staged;
    q.keys = F(w.keys); q.values = d->memory; q.mem_len = d->mem_len;
    q.align = F(w.align) + (size_t)t * Tm; q.s_al = (long long)S * Tm;
    q.dctx = dctx + (size_t)t * D; q.s_dc = (long long)S * D;
    q.datt_next = (last || AL > 0) ? nullptr : F(w.dinp[0]); q.s_dn = D + Ud;  // with the attention layer it is added to datt below
    q.dscore = F(w.dscore) + (size_t)t * Tm; q.s_ds = (long long)S * Tm;
    q.score_p = F(w.psave) + (size_t)t * Tm; q.s_sp = q.s_al;
    q.align_prev = t > 0 ? F(w.align) + (size_t)(t - 1) * Tm : nullptr; q.s_ap = q.s_al;
    q.dalign_in = last ? nullptr : F(w.dalign) + (size_t)((t + 1) & 1) * B * Tm;
    q.dalign_out = F(w.dalign) + (size_t)(t & 1) * B * Tm; q.dbias_acc = F(w.dbias);
    q.dq = F(w.dq); q.s_dq = Ud;
    q.pq = F(w.pq) + (size_t)t * Ud; q.s_pq = sh; q.w_query = d->w_query; q.v_att = d->v_att;
    q.dpq = F(w.dpq) + (size_t)t * Ud; q.s_dpq = sh; q.dkeys = F(w.dkeys); q.dv_acc = F(w.dv_acc);
    for (int e = 0; e < 4; ++e) { q.extra[e] = nullptr; q.s_extra[e] = 0; }
    q.next_base = (long long)(t + 1) * D; q.seed = d->drop_seed; q.thresh = thresh; q.inv_keep = inv_keep; q.step_ptr = d->drop_step;
    {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(B, ns);
      cfg.blockDim = dim3(256);
      cfg.dynamicSmemBytes = attb_smem;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 1;
      attr[0].val.clusterDim.y = (unsigned)ns;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      PLAS_CUDA(cudaLaunchKernelEx(&cfg, dec_att_bwd_kernel, q));
    }
    for (int l = L - 1; l >= 0; --l) {
      const int Kin = (l == 0 ? A : Ud);
      CellBwdArgs c;
      c.B = B; c.Ud = Ud;
      c.z = F(w.z[l]) + (size_t)t * 4 * Ud; c.s_z = sz;
      c.c_t = F(w.c[l]) + (size_t)t * Ud; c.c_prev = t > 0 ? F(w.c[l]) + (size_t)(t - 1) * Ud : nullptr; c.s_c = sh; c.s_cp = sh;
      c.dc_carry = F(w.dc[l]); c.first = last ? 1 : 0;
      c.dq = (l == L - 1) ? F(w.dq) : nullptr; c.s_dq = Ud;
      c.dh_next = last ? nullptr : F(w.dinp[l]) + Kin; c.s_dn = Kin + Ud;
      c.dh_above = (l < L - 1) ? F(w.dinp[l + 1]) : nullptr; c.s_da = 2 * Ud;
      c.dh_extra = (l == L - 1 && AL > 0) ? F(w.dqx) : nullptr; c.s_dx = Ud;  // top layer: the attention layer's h part
      c.idx_base = (long long)t * Ud; c.seed = d->drop_seed + 1 + l; c.thresh = thresh; c.inv_keep = inv_keep; c.step_ptr = d->drop_step;
      dec_cell_bwd_kernel<<<(B * Ud + 255) / 256, 256, 0, st>>>(c);
      GemvTArgs g;
      g.B = B; g.N = 4 * Ud; g.K = Kin + Ud;
      g.dz = c.z; g.s_z = sz;
      g.w = d->kernel[l] + (size_t)(l == 0 ? E : 0) * 4 * Ud;
      g.dinp = F(w.dinp[l]); g.s_o = Kin + Ud;
      {
        const int per = (g.N / 4 + GV_NS - 1) / GV_NS;
        const size_t gsm = ((size_t)DT_ROWS * (4 * per + 4) + (size_t)GV_KG * 4 * per + (size_t)GV_NS * DT_ROWS * 8) * 4;
        PLAS_REQUIRE(gsm <= 220 * 1024, "dec_train_bwd: Ud = %d too large", Ud);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(GV_NS, (g.K + GV_KG - 1) / GV_KG, (B + DT_ROWS - 1) / DT_ROWS);
        cfg.blockDim = dim3(256);
        cfg.dynamicSmemBytes = gsm;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = GV_NS;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        PLAS_CUDA(cudaLaunchKernelEx(&cfg, dec_gemv_t_kernel, g));
      }
    }
  }
  PLAS_CUDA(cudaGetLastError());
  // weight gradients over the B*S saved rows (z now holds dz)
  for (int l = 0; l < L; ++l) {
    PLAS_REQUIRE(d->dkernel[l] && d->dbias[l], "dec_train_bwd: null gradient tensor (layer %d)", l);
    const float* dz = F(w.z[l]);
    float* dk = d->dkernel[l];
    if (l == 0) {
      if ((rc = gemm(st, E, 4 * Ud, (int)BS, d->x_in, 1, E, dz, 4 * Ud, 1, dk, 4 * Ud, nullptr, 0.f, 1, 0, 0, 0, sk, skb))) return rc;
      // gradient wrt the decoder inputs (embedding_size != 0: flows into target_embedding): dX = dZ_0 W_0[0:E]^T
      if (d->dx_in && (rc = gemm(st, BS, E, 4 * Ud, dz, 4 * Ud, 1, d->kernel[0], 1, 4 * Ud, d->dx_in, E))) return rc;
      if ((rc = gemm(st, A, 4 * Ud, (int)BS, F(w.att_prev), 1, A, dz, 4 * Ud, 1, dk + (size_t)E * 4 * Ud, 4 * Ud, nullptr, 0.f, 1, 0, 0, 0, sk, skb))) return rc;
      if ((rc = gemm(st, Ud, 4 * Ud, (int)BS, F(w.hprev[0]), 1, Ud, dz, 4 * Ud, 1, dk + (size_t)(E + A) * 4 * Ud, 4 * Ud, nullptr, 0.f, 1, 0, 0, 0, sk, skb))) return rc;
    } else {
      if ((rc = gemm(st, Ud, 4 * Ud, (int)BS, F(drop ? w.hdrop[l - 1] : w.h[l - 1]), 1, Ud, dz, 4 * Ud, 1, dk, 4 * Ud, nullptr, 0.f, 1, 0, 0, 0, sk, skb))) return rc;
      if ((rc = gemm(st, Ud, 4 * Ud, (int)BS, F(w.hprev[l]), 1, Ud, dz, 4 * Ud, 1, dk + (size_t)Ud * 4 * Ud, 4 * Ud, nullptr, 0.f, 1, 0, 0, 0, sk, skb))) return rc;
    }
    if ((rc = plas_colsum_f32(dz, BS, 4 * Ud, 4 * Ud, d->dbias[l], 0, st))) return rc;
  }
  const float* htop = F(w.h[L - 1]);
  if (AL > 0) {  // dW_att = [H_top; Ctx]^T dAtt
    if ((rc = gemm(st, Ud, A, (int)BS, htop, 1, Ud, datt, A, 1, d->dw_att_layer, A, nullptr, 0.f, 1, 0, 0, 0, sk, skb))) return rc;
    if ((rc = gemm(st, D, A, (int)BS, F(w.ctx), 1, D, datt, A, 1, d->dw_att_layer + (size_t)Ud * A, A, nullptr, 0.f, 1, 0, 0, 0, sk, skb))) return rc;
  }
  if (bah) {
    if ((rc = gemm(st, Ud, Ud, (int)BS, htop, 1, Ud, F(w.dpq), Ud, 1, d->dw_query, Ud))) return rc;
    if ((rc = plas_colsum_f32(F(w.dv_acc), B, Ud, Ud, d->dv_att, 0, st))) return rc;
  } else {
    // dkeys[b] = dScore[b]^T Q[b], Q = the query the score read: H_top, or relu(query_layer(H_top)) (custom, saved in pq)
    if ((rc = gemm(st, Tm, Ud, S, F(w.dscore), 1, Tm, custom ? F(w.pq) : htop, Ud, 1, F(w.dkeys), Ud, nullptr, 0.f, B, (long long)S * Tm,
                   (long long)S * Ud, (long long)Tm * Ud)))
      return rc;
    if (custom) {  // keys = relu(memory_layer(values)); dW_query = H_top^T dPQ
      dec_relu_grad_kernel<<<(unsigned)(((size_t)B * Tm * Ud + 255) / 256), 256, 0, st>>>(F(w.dkeys), F(w.keys), (size_t)B * Tm * Ud);
      PLAS_REQUIRE(d->dw_query != nullptr, "dec_train_bwd: custom attention needs dw_query");
      if ((rc = gemm(st, Ud, Ud, (int)BS, htop, 1, Ud, F(w.dpq), Ud, 1, d->dw_query, Ud))) return rc;
    }
  }
  // dvalues[b] = Align[b]^T dAtt[b]  (accumulated into the encoder-output gradient)
  if ((rc = gemm(st, Tm, D, S, F(w.align), 1, Tm, dctx, D, 1, d->dmemory, D, nullptr, d->dmemory_accumulate ? 1.f : 0.f, B,
                 (long long)S * Tm, (long long)S * D, (long long)Tm * D)))
    return rc;
  // memory_layer: dmemory += dkeys W_mem^T, dW_mem = memory^T dkeys
  if (mono && (rc = plas_colsum_f32(F(w.dbias), B, 1, 1, d->dscore_bias, 0, st))) return rc;
  if ((rc = gemm(st, (long long)B * Tm, D, Ud, F(w.dkeys), Ud, 1, d->w_mem, 1, Ud, d->dmemory, D, nullptr, 1.f))) return rc;
  return gemm(st, D, Ud, B * Tm, d->memory, 1, D, F(w.dkeys), Ud, 1, d->dw_mem, Ud, nullptr, 0.f, 1, 0, 0, 0, sk, skb);
}
