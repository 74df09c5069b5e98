// K3 (training) -- persistent LSTM recurrence forward-with-save and its backward (BPTT) in exact fp32, directly on
// the TF checkpoint layout.  Replaces the forward and the gradient of tf.nn.bidirectional_dynamic_rnn over
// tf.nn.rnn_cell.LSTMCell (reference las/ops.py:23-46; gate order i|j|f|o in column blocks, forget_bias 1.0) for
// the TRAIN graph of model_helper.py:403-417.
//
// Same decomposition as the inference kernel (rec.cu): a (direction, 16-utterance group) is an independent
// recurrence owned by G = U/upc co-resident CTAs that exchange the per-step vector through an L2-resident
// ping-pong buffer with acquire/release counters; W_hh stays in shared memory for the whole sequence.
//   forward : z[b,t] (x-projections + bias, K2) is overwritten IN PLACE by the activated gates (i, tanh j, f, o);
//             c_t and h_{s-1} are saved for the backward pass, h_t goes to the layer output.
//   backward: walks the steps in reverse;  dh_s = dout[t] + dz_{s+1} W_hh^T  (the CTA holds the rows of W_hh of its
//             own units, so the exchanged vector is dz, 4U wide);  gate derivatives from the saved activations;
//             dz_s overwrites the saved gates IN PLACE and is what the weight / input gradient GEMMs consume
//             (dW = [x;h_{s-1}]^T dz, dx = dz W_x^T: plas_gemm_f32_ex).  Positions t >= len are zeroed.
#include "common.cuh"
#include "../../include/plas.h"

namespace plas {

constexpr int RT_THREADS = 256;
constexpr int RT_ROWS = 16;
constexpr int RT_MAXR = 2;

struct RecTrainArgs {
  plas_rec_train_desc d;
  float* xbuf;         // [2][ndir][Bpad][W] exchange (W = U forward, 4U backward)
  unsigned* counters;  // [ndir][n_groups]
  int n_groups, group_offset, groups_here, upc, G, Bpad;
};

__device__ __forceinline__ void rt_wait(const unsigned* ctr, unsigned target) {
  unsigned spins = 0;
  while (ld_acquire_u32(ctr) < target)
    if (++spins > (1u << 28)) __trap();
}

__global__ void __launch_bounds__(RT_THREADS, 1) rec_train_fwd_kernel(RecTrainArgs p) {
  extern __shared__ __align__(16) unsigned char rt_smem[];
  const plas_rec_train_desc& d = p.d;
  const int U = d.U, B = d.B, T = d.T, ndir = d.ndir, upc = p.upc;
  const int tid = threadIdx.x;
  int bid = blockIdx.x;
  const int ci = bid % p.G; bid /= p.G;
  const int gi = p.group_offset + bid % p.groups_here;
  const int dir = bid / p.groups_here;
  const int row0 = gi * RT_ROWS;

  float4* s_w = reinterpret_cast<float4*>(rt_smem);               // [U (k)][upc] : (i,j,f,o) columns of a unit
  float* s_h = reinterpret_cast<float*>(s_w + (size_t)U * upc);  // [16][U]
  __shared__ int s_len[RT_ROWS];
  __shared__ int s_tmax;
  {
    const float* kern = d.kernel[dir] + (size_t)d.din * 4 * U;  // W_hh rows of the TF kernel
    for (int i = tid; i < U * upc; i += RT_THREADS) {
      const int k = i / upc, ul = i % upc;
      const float* row = kern + (size_t)k * 4 * U + ci * upc + ul;
      s_w[i] = make_float4(row[0], row[U], row[2 * U], row[3 * U]);
    }
  }
  if (tid < RT_ROWS) s_len[tid] = (row0 + tid < B) ? min(d.lengths[row0 + tid], T) : 0;
  __syncthreads();
  if (tid == 0) {
    int m = 0;
    for (int r = 0; r < RT_ROWS; ++r) m = max(m, s_len[r]);
    s_tmax = m;
  }
  __syncthreads();
  const int Tg = s_tmax;

  float* hx = p.xbuf;
  unsigned* ctr = p.counters + dir * p.n_groups + gi;
  const int ul = tid % upc, rg = tid / upc, nrg = RT_THREADS / upc;
  const int unit = ci * upc + ul;
  const size_t zrow = (size_t)ndir * 4 * U;  // floats per (b,t) in z
  const size_t srow = (size_t)ndir * U;      // floats per (b,t) in c_save / h_prev
  float c_state[RT_MAXR] = {0.f, 0.f}, h_state[RT_MAXR] = {0.f, 0.f};

  for (int s = 0; s < Tg; ++s) {
    float4 acc[RT_MAXR];
    bool act[RT_MAXR];
    int t_idx[RT_MAXR];
#pragma unroll
    for (int e = 0; e < RT_MAXR; ++e) {
      const int r = rg + e * nrg;
      act[e] = false;
      acc[e] = make_float4(0.f, 0.f, 0.f, 0.f);
      t_idx[e] = 0;
      if (r < RT_ROWS) {
        const int len = s_len[r];
        act[e] = s < len;
        t_idx[e] = dir ? (len - 1 - s) : s;
        if (act[e]) {
          const float* zp = d.z + ((size_t)(row0 + r) * T + t_idx[e]) * zrow + (size_t)dir * 4 * U + unit;
          acc[e] = make_float4(zp[0], zp[U], zp[2 * U], zp[3 * U]);
        }
      }
    }
    if (s > 0) {
      if (tid == 0) rt_wait(ctr, (unsigned)(p.G * s));
      __syncthreads();
      const float4* hsrc = reinterpret_cast<const float4*>(hx + (((size_t)((s - 1) & 1) * ndir + dir) * p.Bpad + row0) * U);
      float4* hdst = reinterpret_cast<float4*>(s_h);
      for (int i = tid; i < RT_ROWS * U / 4; i += RT_THREADS) hdst[i] = __ldcg(hsrc + i);
      __syncthreads();
      for (int k = 0; k < U; ++k) {
        const float4 w = s_w[(size_t)k * upc + ul];
#pragma unroll
        for (int e = 0; e < RT_MAXR; ++e) {
          const int r = rg + e * nrg;
          if (r < RT_ROWS) {
            const float hv = s_h[r * U + k];
            acc[e].x = fmaf(hv, w.x, acc[e].x);
            acc[e].y = fmaf(hv, w.y, acc[e].y);
            acc[e].z = fmaf(hv, w.z, acc[e].z);
            acc[e].w = fmaf(hv, w.w, acc[e].w);
          }
        }
      }
    }
#pragma unroll
    for (int e = 0; e < RT_MAXR; ++e) {
      const int r = rg + e * nrg;
      if (r >= RT_ROWS) continue;
      const int b = row0 + r;
      if (act[e]) {
        const float gi_ = sigmoidf_acc(acc[e].x), gj = tanhf(acc[e].y), gf = sigmoidf_acc(acc[e].z + 1.0f),
                    go = sigmoidf_acc(acc[e].w);
        const float cn = gf * c_state[e] + gi_ * gj;
        const float hn = go * tanhf(cn);
        const size_t bt = (size_t)b * T + t_idx[e];
        float* zp = d.z + bt * zrow + (size_t)dir * 4 * U + unit;
        zp[0] = gi_; zp[U] = gj; zp[2 * U] = gf; zp[3 * U] = go;
        d.c_save[bt * srow + dir * U + unit] = cn;
        d.h_prev[bt * srow + dir * U + unit] = h_state[e];
        d.out[(size_t)b * d.out_batch_stride + (size_t)t_idx[e] * srow + dir * U + unit] = hn;
        c_state[e] = cn;
        h_state[e] = hn;
      }
      if (b < p.Bpad) hx[(((size_t)(s & 1) * ndir + dir) * p.Bpad + b) * U + unit] = h_state[e];
    }
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      red_release_add_u32(ctr, 1u);
    }
  }
}

__global__ void __launch_bounds__(RT_THREADS, 1) rec_train_bwd_kernel(RecTrainArgs p) {
  extern __shared__ __align__(16) unsigned char rt_smem[];
  const plas_rec_train_desc& d = p.d;
  const int U = d.U, B = d.B, T = d.T, ndir = d.ndir, upc = p.upc;
  const int tid = threadIdx.x;
  int bid = blockIdx.x;
  const int ci = bid % p.G; bid /= p.G;
  const int gi = p.group_offset + bid % p.groups_here;
  const int dir = bid / p.groups_here;
  const int row0 = gi * RT_ROWS;

  float4* s_w = reinterpret_cast<float4*>(rt_smem);                  // [U (n4)][upc]: W_hh[unit][4*n4 .. 4*n4+3]
  float4* s_dz = s_w + (size_t)U * upc;                              // [16][U] float4 = [16][4U]
  __shared__ int s_len[RT_ROWS];
  __shared__ int s_tmax;
  {
    const float4* kern = reinterpret_cast<const float4*>(d.kernel[dir] + (size_t)d.din * 4 * U);
    for (int i = tid; i < U * upc; i += RT_THREADS) {
      const int ul = i / U, n4 = i % U;
      s_w[(size_t)n4 * upc + ul] = kern[(size_t)(ci * upc + ul) * U + n4];
    }
  }
  if (tid < RT_ROWS) s_len[tid] = (row0 + tid < B) ? min(d.lengths[row0 + tid], T) : 0;
  __syncthreads();
  if (tid == 0) {
    int m = 0;
    for (int r = 0; r < RT_ROWS; ++r) m = max(m, s_len[r]);
    s_tmax = m;
  }
  __syncthreads();
  const int Tg = s_tmax;

  float* dzx = p.xbuf;
  unsigned* ctr = p.counters + dir * p.n_groups + gi;
  const int ul = tid % upc, rg = tid / upc, nrg = RT_THREADS / upc;
  const int unit = ci * upc + ul;
  const size_t zrow = (size_t)ndir * 4 * U;
  const size_t srow = (size_t)ndir * U;
  const int W4 = 4 * U;

  // the gradient GEMMs run over every (b,t) row: zero dz past each utterance's length
#pragma unroll
  for (int e = 0; e < RT_MAXR; ++e) {
    const int r = rg + e * nrg;
    if (r >= RT_ROWS || row0 + r >= B) continue;
    for (int t = s_len[r]; t < T; ++t) {
      float* zp = d.z + ((size_t)(row0 + r) * T + t) * zrow + (size_t)dir * 4 * U + unit;
      zp[0] = 0.f; zp[U] = 0.f; zp[2 * U] = 0.f; zp[3 * U] = 0.f;
    }
  }

  float dc_carry[RT_MAXR] = {0.f, 0.f};
  for (int j = 0; j < Tg; ++j) {
    const int s = Tg - 1 - j;
    bool act[RT_MAXR];
    int t_idx[RT_MAXR];
    float4 g[RT_MAXR];
    float c_t[RT_MAXR], c_prev[RT_MAXR], dh[RT_MAXR];
#pragma unroll
    for (int e = 0; e < RT_MAXR; ++e) {
      const int r = rg + e * nrg;
      act[e] = false;
      t_idx[e] = 0;
      dh[e] = 0.f;
      c_t[e] = c_prev[e] = 0.f;
      g[e] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < RT_ROWS) {
        const int len = s_len[r];
        act[e] = s < len;
        t_idx[e] = dir ? (len - 1 - s) : s;
        if (act[e]) {
          const int b = row0 + r;
          const size_t bt = (size_t)b * T + t_idx[e];
          const float* zp = d.z + bt * zrow + (size_t)dir * 4 * U + unit;
          g[e] = make_float4(zp[0], zp[U], zp[2 * U], zp[3 * U]);
          c_t[e] = d.c_save[bt * srow + dir * U + unit];
          if (s > 0) {
            const size_t btp = (size_t)b * T + (dir ? t_idx[e] + 1 : t_idx[e] - 1);
            c_prev[e] = d.c_save[btp * srow + dir * U + unit];
          }
          dh[e] = d.dout[(size_t)b * d.out_batch_stride + (size_t)t_idx[e] * srow + dir * U + unit];
        }
      }
    }
    if (j > 0) {
      if (tid == 0) rt_wait(ctr, (unsigned)(p.G * j));
      __syncthreads();
      const float4* src = reinterpret_cast<const float4*>(dzx + (((size_t)((j - 1) & 1) * ndir + dir) * p.Bpad + row0) * W4);
      for (int i = tid; i < RT_ROWS * U; i += RT_THREADS) s_dz[i] = __ldcg(src + i);
      __syncthreads();
      float racc[RT_MAXR] = {0.f, 0.f};
      for (int n4 = 0; n4 < U; ++n4) {
        const float4 w = s_w[(size_t)n4 * upc + ul];
#pragma unroll
        for (int e = 0; e < RT_MAXR; ++e) {
          const int r = rg + e * nrg;
          if (r < RT_ROWS) {
            const float4 v = s_dz[r * U + n4];
            racc[e] = fmaf(v.x, w.x, racc[e]);
            racc[e] = fmaf(v.y, w.y, racc[e]);
            racc[e] = fmaf(v.z, w.z, racc[e]);
            racc[e] = fmaf(v.w, w.w, racc[e]);
          }
        }
      }
#pragma unroll
      for (int e = 0; e < RT_MAXR; ++e) dh[e] += racc[e];
    }
#pragma unroll
    for (int e = 0; e < RT_MAXR; ++e) {
      const int r = rg + e * nrg;
      if (r >= RT_ROWS) continue;
      const int b = row0 + r;
      float dzi = 0.f, dzj = 0.f, dzf = 0.f, dzo = 0.f;
      if (act[e]) {
        const float gi_ = g[e].x, gj = g[e].y, gf = g[e].z, go = g[e].w;
        const float tc = tanhf(c_t[e]);
        const float dc = dc_carry[e] + dh[e] * go * (1.f - tc * tc);
        dzo = dh[e] * tc * go * (1.f - go);
        dzi = dc * gj * gi_ * (1.f - gi_);
        dzj = dc * gi_ * (1.f - gj * gj);
        dzf = dc * c_prev[e] * gf * (1.f - gf);
        dc_carry[e] = dc * gf;
        float* zp = d.z + ((size_t)b * T + t_idx[e]) * zrow + (size_t)dir * 4 * U + unit;
        zp[0] = dzi; zp[U] = dzj; zp[2 * U] = dzf; zp[3 * U] = dzo;
      }
      if (b < p.Bpad) {
        float* xp = dzx + (((size_t)(j & 1) * ndir + dir) * p.Bpad + b) * W4 + unit;
        xp[0] = dzi; xp[U] = dzj; xp[2 * U] = dzf; xp[3 * U] = dzo;
      }
    }
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      red_release_add_u32(ctr, 1u);
    }
  }
}

static int rt_upc(int U) {
  int upc = 32;
  while (upc > 4 && ((size_t)U * 16 * upc > 128 * 1024 || U / upc < 8)) upc >>= 1;  // >= 8 CTAs per group when possible
  while (upc > 4 && U % upc != 0) upc >>= 1;
  return upc;
}

static void rt_ws_layout(const plas_rec_train_desc& d, size_t* o_ctr, size_t* o_x, size_t* total) {
  const int n_groups = (d.B + RT_ROWS - 1) / RT_ROWS;
  size_t off = 0;
  *o_ctr = off;
  off += ((size_t)d.ndir * n_groups * 4 + 255) & ~size_t(255);
  *o_x = off;
  off += ((size_t)2 * d.ndir * n_groups * RT_ROWS * 4 * d.U * 4 + 255) & ~size_t(255);
  *total = off;
}

static int rt_launch(const plas_rec_train_desc* d, void* workspace, size_t workspace_bytes, cudaStream_t stream,
                     bool backward) {
  PLAS_REQUIRE(d && workspace, "rec_train: null argument");
  PLAS_REQUIRE(d->B > 0 && d->T > 0 && d->U > 0 && (d->ndir == 1 || d->ndir == 2) && d->din > 0, "rec_train: bad shape");
  PLAS_REQUIRE(d->z && d->kernel[0] && (d->ndir == 1 || d->kernel[1]) && d->lengths && d->c_save, "rec_train: null tensor");
  PLAS_REQUIRE(backward ? d->dout != nullptr : (d->out != nullptr && d->h_prev != nullptr), "rec_train: null tensor");
  PLAS_REQUIRE(d->U % 4 == 0 && d->U <= 1024, "rec_train: U=%d unsupported", d->U);
  const int upc = rt_upc(d->U);
  PLAS_REQUIRE(d->U % upc == 0 && RT_THREADS % upc == 0 && RT_THREADS / upc * RT_MAXR >= RT_ROWS,
               "rec_train: U=%d unsupported (upc=%d)", d->U, upc);
  PLAS_REQUIRE(d->out_batch_stride >= (int64_t)d->T * d->ndir * d->U, "rec_train: out_batch_stride too small");
  size_t o_ctr, o_x, total;
  rt_ws_layout(*d, &o_ctr, &o_x, &total);
  PLAS_REQUIRE(workspace_bytes >= total, "rec_train: workspace %zu < %zu", workspace_bytes, total);
  RecTrainArgs a;
  a.d = *d;
  a.counters = (unsigned*)((unsigned char*)workspace + o_ctr);
  a.xbuf = (float*)((unsigned char*)workspace + o_x);
  a.n_groups = (d->B + RT_ROWS - 1) / RT_ROWS;
  a.upc = upc;
  a.G = d->U / upc;
  a.Bpad = a.n_groups * RT_ROWS;
  PLAS_CUDA(cudaMemsetAsync(a.counters, 0, (size_t)d->ndir * a.n_groups * 4, stream));
  const size_t smem = (size_t)d->U * upc * 16 + (size_t)RT_ROWS * d->U * 4 * (backward ? 4 : 1);
  PLAS_REQUIRE(smem <= 227 * 1024, "rec_train: needs %zu bytes of shared memory", smem);
  const void* fn = backward ? (const void*)rec_train_bwd_kernel : (const void*)rec_train_fwd_kernel;
  PLAS_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  PLAS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, RT_THREADS, smem));
  const int resident = per_sm * num_sms();
  const int ctas_per_group = d->ndir * a.G;
  const int gpl = resident / ctas_per_group;
  PLAS_REQUIRE(gpl >= 1, "rec_train: %d CTAs per group cannot be co-resident (%d slots)", ctas_per_group, resident);
  for (int g0 = 0; g0 < a.n_groups; g0 += gpl) {
    a.group_offset = g0;
    a.groups_here = (a.n_groups - g0 < gpl) ? (a.n_groups - g0) : gpl;
    void* args[] = {&a};
    PLAS_CUDA(cudaLaunchCooperativeKernel(fn, dim3(ctas_per_group * a.groups_here), dim3(RT_THREADS), args, smem, stream));
  }
  return PLAS_OK;
}

}  // namespace plas

using namespace plas;

extern "C" size_t plas_rec_train_workspace_bytes(const plas_rec_train_desc* d) {
  size_t a, b, total;
  rt_ws_layout(*d, &a, &b, &total);
  return total;
}

extern "C" int plas_bilstm_rec_train_fwd(const plas_rec_train_desc* d, void* workspace, size_t workspace_bytes,
                                         plas_stream_t stream) {
  return rt_launch(d, workspace, workspace_bytes, (cudaStream_t)stream, false);
}

extern "C" int plas_bilstm_rec_train_bwd(const plas_rec_train_desc* d, void* workspace, size_t workspace_bytes,
                                         plas_stream_t stream) {
  return rt_launch(d, workspace, workspace_bytes, (cudaStream_t)stream, true);
}
