// K3 (training) -- persistent LSTM recurrence forward-with-save and its backward (BPTT) in exact fp32, directly on
// the TF checkpoint layout.  Replaces the forward and the gradient of tf.nn.bidirectional_dynamic_rnn over
// tf.nn.rnn_cell.LSTMCell (reference las/ops.py:23-46; gate order i|j|f|o in column blocks, forget_bias 1.0) for
// the TRAIN graph of model_helper.py:403-417.
//
// Decomposition (as in the inference kernel rec.cu): a (direction, group of R utterances) is an independent
// recurrence owned by G = U/UPC co-resident CTAs (cooperative launch) that exchange the per-step vector through an
// L2-resident ping-pong buffer with acquire/release counters; the CTA's slice of W_hh stays in shared memory for
// the whole sequence.  256 threads = UPC units x R rows x KS slices of the reduction (partials summed in a fixed
// order through shared memory).
//   forward : z[b,t] (x-projections + bias, K2) is overwritten IN PLACE by the activated gates (i, tanh j, f, o);
//             c_t and h_{s-1} are saved for the backward pass, h_t goes to the layer output.  Exchanged: h (U wide).
//   backward: walks the steps in reverse;  dh_s = dout[t] + dz_{s+1} W_hh^T  (the CTA holds the ROWS of W_hh of its
//             own units, so the exchanged vector is dz, 4U wide -- hence small row groups, R = 8);  gate derivatives
//             from the saved activations;  dz_s overwrites the saved gates IN PLACE and is what the weight / input
//             gradient GEMMs consume (dW = [x;h_{s-1}]^T dz, dx = dz W_x^T: plas_gemm_f32_ex).  t >= len is zeroed.
// The (R, UPC, RB) shape is picked per call so that every group's CTAs are co-resident in as few launches as possible.
// Forward only, when U/32 <= 8 and every (direction, group) fits in one wave: the CTAs of a group form a thread-block
// cluster and exchange h through distributed shared memory with one cluster barrier per step (2.57 ms vs 2.86 ms at c3).
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"
#include "../../include/plas.h"

namespace plas {

constexpr int RT_THREADS = 256;

struct RecTrainArgs {
  plas_rec_train_desc d;
  float* xbuf;         // [2][ndir][Bpad][W] exchange (W = U forward, 4U backward)
  unsigned* counters;  // [ndir][n_groups]
  int n_groups, group_offset, groups_here, G, Bpad;
};

__device__ __forceinline__ void rt_wait(const unsigned* ctr, unsigned target) {
  unsigned spins = 0;
  while (ld_acquire_u32(ctr) < target)
    if (++spins > (1u << 28)) __trap();
}

// Thread = (unit ul, block of RB rows, k-slice kh): one weight vector fetched from shared memory feeds RB rows from
// registers, so the W_hh slice is read R/RB times per step instead of R times (with RB = 1 the step is bound by those
// shared-memory reads: ncu short_scoreboard + barrier stalls, FMA pipe 13 % active).
// CL: the G <= 8 CTAs of a (direction, group) form one thread-block cluster; h_s is pushed in 16-byte pieces into the
// double-buffered tile of every CTA through distributed shared memory and ONE cluster barrier per step replaces the three
// L2 round trips of the counter protocol (poll, tile load, fence + atomic).
template <int R, int UPC, int RB, bool CL>
__global__ void __launch_bounds__(RT_THREADS, 1) rec_train_fwd_kernel(RecTrainArgs p) {
  namespace cg = cooperative_groups;
  constexpr int NRB = R / RB, TPS = UPC * NRB, KS = RT_THREADS / TPS;
  static_assert(R % RB == 0 && KS >= 1 && KS * TPS == RT_THREADS && R * UPC <= RT_THREADS, "bad shape");
  extern __shared__ __align__(16) unsigned char rt_smem[];
  const plas_rec_train_desc& d = p.d;
  const int U = d.U, B = d.B, T = d.T, ndir = d.ndir;
  const int tid = threadIdx.x;
  int bid = blockIdx.x;
  const int ci = bid % p.G; bid /= p.G;
  const int gi = p.group_offset + bid % p.groups_here;
  const int dir = bid / p.groups_here;
  const int row0 = gi * R;
  const int HS = U + 4;  // padded row stride of the staged h tile

  float4* s_w = reinterpret_cast<float4*>(rt_smem);               // [U (k)][UPC] : (i,j,f,o) columns of a unit
  float* s_h = reinterpret_cast<float*>(s_w + (size_t)U * UPC);  // [2 if CL][R][HS]
  float4* s_part = reinterpret_cast<float4*>(s_h + (size_t)(CL ? 2 : 1) * R * HS);  // [KS][R][UPC] partial pre-activations
  float* s_stage = reinterpret_cast<float*>(s_part + (size_t)KS * R * UPC);          // CL: [R][UPC] slice to push
  __shared__ int s_len[R];
  __shared__ int s_tmax;
  {
    const float* kern = d.kernel[dir] + (size_t)d.din * 4 * U;  // W_hh rows of the TF kernel
    for (int i = tid; i < U * UPC; i += RT_THREADS) {
      const int k = i / UPC, ul = i % UPC;
      const float* row = kern + (size_t)k * 4 * U + ci * UPC + ul;
      s_w[i] = make_float4(row[0], row[U], row[2 * U], row[3 * U]);
    }
  }
  if (tid < R) s_len[tid] = (row0 + tid < B) ? min(d.lengths[row0 + tid], T) : 0;
  __syncthreads();
  if (tid == 0) {
    int m = 0;
    for (int r = 0; r < R; ++r) m = max(m, s_len[r]);
    s_tmax = m;
  }
  __syncthreads();
  const int Tg = s_tmax;

  float* hx = p.xbuf;
  unsigned* ctr = p.counters + dir * p.n_groups + gi;
  // compute role: (unit ul, row block r0.., k-slice kh);  finalise role (tid < R*UPC): one (row fr, unit ful) pair --
  // the gate math is spread over R*UPC threads instead of the few kh == 0 threads
  const int ul = tid % UPC, r0 = ((tid / UPC) % NRB) * RB, kh = tid / TPS;
  const bool fin = tid < R * UPC;
  const int fr = fin ? tid / UPC : 0, ful = tid % UPC;
  const int unit = ci * UPC + ful;
  const int b = row0 + fr;
  const int len = s_len[fr];
  const size_t zrow = (size_t)ndir * 4 * U;
  const size_t srow = (size_t)ndir * U;
  const int kper = U / KS;  // host guarantees (U / KS) % 4 == 0
  float c_state = 0.f, h_state = 0.f;
  if (CL) cg::this_cluster().sync();  // every CTA of the cluster is running before its shared memory is written remotely

  for (int s = 0; s < Tg; ++s) {
    const bool act = fin && s < len;
    const int t_idx = dir ? (len - 1 - s) : s;
    float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (act) {  // prefetch this step's pre-activations; they are consumed after the exchange
      const float* zp = d.z + ((size_t)b * T + t_idx) * zrow + (size_t)dir * 4 * U + unit;
      z = make_float4(zp[0], zp[U], zp[2 * U], zp[3 * U]);
    }
    if (s > 0) {
      if (!CL) {
        if (tid == 0) rt_wait(ctr, (unsigned)(p.G * s));
        __syncthreads();
        const float4* hsrc = reinterpret_cast<const float4*>(hx + (((size_t)((s - 1) & 1) * ndir + dir) * p.Bpad + row0) * U);
        const int U4 = U / 4;
        for (int i = tid; i < R * U4; i += RT_THREADS) {
          const int rr = i / U4, c4 = i - rr * U4;
          *reinterpret_cast<float4*>(s_h + rr * HS + 4 * c4) = __ldcg(hsrc + i);
        }
        __syncthreads();
      }
      float4 acc[RB];
#pragma unroll
      for (int i = 0; i < RB; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* hrow = s_h + (CL ? (size_t)((s - 1) & 1) * R * HS : 0) + r0 * HS + kh * kper;
      const float4* wcol = s_w + (size_t)(kh * kper) * UPC + ul;
#pragma unroll 2
      for (int k = 0; k < kper; k += 4) {
        const float4 w0 = wcol[(size_t)(k + 0) * UPC], w1 = wcol[(size_t)(k + 1) * UPC];
        const float4 w2 = wcol[(size_t)(k + 2) * UPC], w3 = wcol[(size_t)(k + 3) * UPC];
#pragma unroll
        for (int i = 0; i < RB; ++i) {
          const float4 hv = *reinterpret_cast<const float4*>(hrow + i * HS + k);
          float4& a = acc[i];
          a.x = fmaf(hv.x, w0.x, a.x); a.y = fmaf(hv.x, w0.y, a.y); a.z = fmaf(hv.x, w0.z, a.z); a.w = fmaf(hv.x, w0.w, a.w);
          a.x = fmaf(hv.y, w1.x, a.x); a.y = fmaf(hv.y, w1.y, a.y); a.z = fmaf(hv.y, w1.z, a.z); a.w = fmaf(hv.y, w1.w, a.w);
          a.x = fmaf(hv.z, w2.x, a.x); a.y = fmaf(hv.z, w2.y, a.y); a.z = fmaf(hv.z, w2.z, a.z); a.w = fmaf(hv.z, w2.w, a.w);
          a.x = fmaf(hv.w, w3.x, a.x); a.y = fmaf(hv.w, w3.y, a.y); a.z = fmaf(hv.w, w3.z, a.z); a.w = fmaf(hv.w, w3.w, a.w);
        }
      }
#pragma unroll
      for (int i = 0; i < RB; ++i) s_part[(kh * R + r0 + i) * UPC + ul] = acc[i];
      __syncthreads();
      if (fin) {
#pragma unroll
        for (int q = 0; q < KS; ++q) {  // fixed order: deterministic
          const float4 o = s_part[(q * R + fr) * UPC + ful];
          z.x += o.x; z.y += o.y; z.z += o.z; z.w += o.w;
        }
      }
    }
    if (fin) {
      if (act) {
        const float gi_ = sigmoidf_acc(z.x), gj = tanhf(z.y), gf = sigmoidf_acc(z.z + 1.0f), go = sigmoidf_acc(z.w);
        const float cn = gf * c_state + gi_ * gj;
        const float hn = go * tanhf(cn);
        const size_t bt = (size_t)b * T + t_idx;
        float* zp = d.z + bt * zrow + (size_t)dir * 4 * U + unit;
        zp[0] = gi_; zp[U] = gj; zp[2 * U] = gf; zp[3 * U] = go;
        if (d.c_save) d.c_save[bt * srow + dir * U + unit] = cn;   // saved only for a later backward call
        if (d.h_prev) d.h_prev[bt * srow + dir * U + unit] = h_state;
        d.out[(size_t)b * d.out_batch_stride + (size_t)t_idx * srow + dir * U + unit] = hn;
        c_state = cn;
        h_state = hn;
      }
      if (CL) s_stage[fr * UPC + ful] = h_state;
      else hx[(((size_t)(s & 1) * ndir + dir) * p.Bpad + b) * U + unit] = h_state;
    }
    __syncthreads();
    if (CL) {
      constexpr int Q = UPC / 4;  // 16-byte pieces per row of the CTA's slice
      float* tile = s_h + (size_t)(s & 1) * R * HS + ci * UPC;
      for (int i = tid; i < R * Q * p.G; i += RT_THREADS) {
        const int c = i / (R * Q), e = i - c * (R * Q);
        const int rr = e / Q, q = e - rr * Q;
        const float4 v = reinterpret_cast<const float4*>(s_stage)[rr * Q + q];
        *reinterpret_cast<float4*>(cg::this_cluster().map_shared_rank(tile + rr * HS + 4 * q, c)) = v;
      }
      cg::this_cluster().sync();  // h_s has landed everywhere; everyone is done reading h_{s-1}
    } else if (tid == 0) {
      __threadfence();
      red_release_add_u32(ctr, 1u);
    }
  }
  if (fin && b < B && d.c_final) {  // final (c, h) of every utterance: the state is frozen past its length
    d.c_final[((size_t)dir * B + b) * U + unit] = c_state;
    d.h_final[((size_t)dir * B + b) * U + unit] = h_state;
  }
}

// (A cluster/DSMEM exchange was measured for the backward pass too: 4.06 ms vs 3.53 ms for the three layers of c3 -- dz is
// four times wider than h, 32 KB land in every CTA per step -- so the backward recurrence keeps the L2 exchange.)
template <int R, int UPC, int RB>
__global__ void __launch_bounds__(RT_THREADS, 1) rec_train_bwd_kernel(RecTrainArgs p) {
  constexpr int NRB = R / RB, TPS = UPC * NRB, KS = RT_THREADS / TPS;
  static_assert(R % RB == 0 && KS >= 1 && KS * TPS == RT_THREADS && R * UPC <= RT_THREADS, "bad shape");
  extern __shared__ __align__(16) unsigned char rt_smem[];
  const plas_rec_train_desc& d = p.d;
  const int U = d.U, B = d.B, T = d.T, ndir = d.ndir;
  const int tid = threadIdx.x;
  int bid = blockIdx.x;
  const int ci = bid % p.G; bid /= p.G;
  const int gi = p.group_offset + bid % p.groups_here;
  const int dir = bid / p.groups_here;
  const int row0 = gi * R;
  const int ZS = U + 1;  // padded row stride (float4 units) of the staged dz tile

  float4* s_w = reinterpret_cast<float4*>(rt_smem);  // [U (n4)][UPC]: W_hh[unit][4*n4 .. 4*n4+3]
  float4* s_dz = s_w + (size_t)U * UPC;              // [R][ZS] float4 = [R][4U]
  float* s_part = reinterpret_cast<float*>(s_dz + (size_t)R * ZS);  // [KS][R][UPC] partial dh
  __shared__ int s_len[R];
  __shared__ int s_tmax;
  {
    const float4* kern = reinterpret_cast<const float4*>(d.kernel[dir] + (size_t)d.din * 4 * U);
    for (int i = tid; i < U * UPC; i += RT_THREADS) {
      const int ul = i / U, n4 = i % U;
      s_w[(size_t)n4 * UPC + ul] = kern[(size_t)(ci * UPC + ul) * U + n4];
    }
  }
  if (tid < R) s_len[tid] = (row0 + tid < B) ? min(d.lengths[row0 + tid], T) : 0;
  __syncthreads();
  if (tid == 0) {
    int m = 0;
    for (int r = 0; r < R; ++r) m = max(m, s_len[r]);
    s_tmax = m;
  }
  __syncthreads();
  const int Tg = s_tmax;

  float* dzx = p.xbuf;
  unsigned* ctr = p.counters + dir * p.n_groups + gi;
  const int ul = tid % UPC, r0 = ((tid / UPC) % NRB) * RB, kh = tid / TPS;  // compute role
  const bool fin = tid < R * UPC;                                            // finalise role: one (row, unit) pair
  const int fr = fin ? tid / UPC : 0, ful = tid % UPC;
  const int unit = ci * UPC + ful;
  const int b = row0 + fr;
  const int len = s_len[fr];
  const size_t zrow = (size_t)ndir * 4 * U;
  const size_t srow = (size_t)ndir * U;
  const int W4 = 4 * U;
  const int nper = U / KS;  // float4 chunks of the reduction per slice

  // the gradient GEMMs run over every (b,t) row: zero dz past each utterance's length
  if (fin && b < B) {
    for (int t = len; t < T; ++t) {
      float* zp = d.z + ((size_t)b * T + t) * zrow + (size_t)dir * 4 * U + unit;
      zp[0] = 0.f; zp[U] = 0.f; zp[2 * U] = 0.f; zp[3 * U] = 0.f;
    }
  }

  float dc_carry = 0.f;
  for (int j = 0; j < Tg; ++j) {
    const int s = Tg - 1 - j;
    const bool act = fin && s < len;
    const int t_idx = dir ? (len - 1 - s) : s;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    float c_t = 0.f, c_prev = 0.f, dh = 0.f;
    if (act) {
      const size_t bt = (size_t)b * T + t_idx;
      const float* zp = d.z + bt * zrow + (size_t)dir * 4 * U + unit;
      g = make_float4(zp[0], zp[U], zp[2 * U], zp[3 * U]);
      c_t = d.c_save[bt * srow + dir * U + unit];
      if (s > 0) {
        const size_t btp = (size_t)b * T + (dir ? t_idx + 1 : t_idx - 1);
        c_prev = d.c_save[btp * srow + dir * U + unit];
      }
      dh = d.dout[(size_t)b * d.out_batch_stride + (size_t)t_idx * srow + dir * U + unit];
      if (s == len - 1 && d.dh_final) {  // the final state IS the state after the last active step (pass_hidden_state)
        dh += d.dh_final[((size_t)dir * B + b) * U + unit];
        dc_carry += d.dc_final[((size_t)dir * B + b) * U + unit];
      }
    }
    if (j > 0) {
      if (tid == 0) rt_wait(ctr, (unsigned)(p.G * j));
      __syncthreads();
      const float4* src = reinterpret_cast<const float4*>(dzx + (((size_t)((j - 1) & 1) * ndir + dir) * p.Bpad + row0) * W4);
      if ((U & (U - 1)) == 0) {  // power-of-two widths: shift / mask instead of a run-time division per element
        const int lg = 31 - __clz(U);
        for (int i = tid; i < R * U; i += RT_THREADS) s_dz[(i >> lg) * ZS + (i & (U - 1))] = __ldcg(src + i);
      } else {
        for (int i = tid; i < R * U; i += RT_THREADS) {
          const int rr = i / U, c4 = i - rr * U;
          s_dz[rr * ZS + c4] = __ldcg(src + i);
        }
      }
      __syncthreads();
      float racc[RB];
#pragma unroll
      for (int i = 0; i < RB; ++i) racc[i] = 0.f;
      const float4* zr = s_dz + r0 * ZS + kh * nper;
      const float4* wc = s_w + (size_t)(kh * nper) * UPC + ul;
#pragma unroll 4
      for (int n4 = 0; n4 < nper; ++n4) {
        const float4 w = wc[(size_t)n4 * UPC];
#pragma unroll
        for (int i = 0; i < RB; ++i) {
          const float4 v = zr[i * ZS + n4];
          racc[i] = fmaf(v.x, w.x, racc[i]);
          racc[i] = fmaf(v.y, w.y, racc[i]);
          racc[i] = fmaf(v.z, w.z, racc[i]);
          racc[i] = fmaf(v.w, w.w, racc[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < RB; ++i) s_part[(kh * R + r0 + i) * UPC + ul] = racc[i];
      __syncthreads();
      if (fin) {
#pragma unroll
        for (int q = 0; q < KS; ++q) dh += s_part[(q * R + fr) * UPC + ful];  // fixed order: deterministic
      }
    }
    if (fin) {
      float dzi = 0.f, dzj = 0.f, dzf = 0.f, dzo = 0.f;
      if (act) {
        const float gi_ = g.x, gj = g.y, gf = g.z, go = g.w;
        const float tc = tanhf(c_t);
        const float dc = dc_carry + dh * go * (1.f - tc * tc);
        dzo = dh * tc * go * (1.f - go);
        dzi = dc * gj * gi_ * (1.f - gi_);
        dzj = dc * gi_ * (1.f - gj * gj);
        dzf = dc * c_prev * gf * (1.f - gf);
        dc_carry = dc * gf;
        float* zp = d.z + ((size_t)b * T + t_idx) * zrow + (size_t)dir * 4 * U + unit;
        zp[0] = dzi; zp[U] = dzj; zp[2 * U] = dzf; zp[3 * U] = dzo;
      }
      float* xp = dzx + (((size_t)(j & 1) * ndir + dir) * p.Bpad + b) * W4 + unit;
      xp[0] = dzi; xp[U] = dzj; xp[2 * U] = dzf; xp[3 * U] = dzo;
    }
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      red_release_add_u32(ctr, 1u);
    }
  }
}

// ---- host side ---------------------------------------------------------------------------------------------
struct RtShape {
  int R, UPC, RB;
  const void* fn;
};
#define RT_FWD(R, UPC, RB) {R, UPC, RB, (const void*)rec_train_fwd_kernel<R, UPC, RB, false>}
#define RT_BWD(R, UPC, RB) {R, UPC, RB, (const void*)rec_train_bwd_kernel<R, UPC, RB>}
// thread-block-cluster variant of the forward recurrence: 8 utterances per group, 32 units per CTA (G = U/32 <= 8), 4 rows per thread (KS = 4)
constexpr int RT_CL_R = 8, RT_CL_UPC = 32, RT_CL_RB = 4, RT_CL_KS = RT_THREADS / (RT_CL_UPC * RT_CL_R / RT_CL_RB);
// order = preference among shapes needing the same number of launches; the RB = 1 shapes serve narrow layers
static const RtShape rt_fwd_shapes[] = {RT_FWD(16, 8, 4), RT_FWD(32, 4, 4), RT_FWD(16, 8, 1), RT_FWD(16, 4, 1), RT_FWD(32, 4, 1), RT_FWD(8, 32, 1)};
static const RtShape rt_bwd_shapes[] = {RT_BWD(8, 16, 4), RT_BWD(8, 16, 1), RT_BWD(8, 32, 1), RT_BWD(8, 8, 1), RT_BWD(8, 4, 1), RT_BWD(16, 4, 1)};

static size_t rt_smem_bytes(const RtShape& sh, int U, bool backward) {
  const int KS = RT_THREADS / (sh.R / sh.RB * sh.UPC);
  if (backward) return (size_t)U * sh.UPC * 16 + (size_t)sh.R * (U + 1) * 16 + (size_t)KS * sh.R * sh.UPC * 4;
  return (size_t)U * sh.UPC * 16 + (size_t)sh.R * (U + 4) * 4 + (size_t)KS * sh.R * sh.UPC * 16;
}

static bool rt_shape_ok(const RtShape& sh, int U, bool backward) {
  const int KS = RT_THREADS / (sh.R / sh.RB * sh.UPC);
  // forward slices k in float4 steps, backward slices the U float4 chunks of a W_hh row
  return U % sh.UPC == 0 && U % (backward ? KS : 4 * KS) == 0 && KS <= 16 && rt_smem_bytes(sh, U, backward) <= 200 * 1024;
}

static void rt_ws_layout(const plas_rec_train_desc& d, size_t* o_ctr, size_t* o_x, size_t* total) {
  const size_t bpad = ((size_t)d.B + 31) / 32 * 32;  // covers every row-group size
  size_t off = 0;
  *o_ctr = off;
  off += ((size_t)d.ndir * ((d.B + 7) / 8) * 4 + 255) & ~size_t(255);
  *o_x = off;
  off += ((size_t)2 * d.ndir * bpad * 4 * d.U * 4 + 255) & ~size_t(255);
  *total = off;
}

static int rt_launch(const plas_rec_train_desc* d, void* workspace, size_t workspace_bytes, cudaStream_t stream,
                     bool backward) {
  PLAS_REQUIRE(d && workspace, "rec_train: null argument");
  PLAS_REQUIRE(d->B > 0 && d->T > 0 && d->U > 0 && (d->ndir == 1 || d->ndir == 2) && d->din > 0, "rec_train: bad shape");
  PLAS_REQUIRE(d->z && d->kernel[0] && (d->ndir == 1 || d->kernel[1]) && d->lengths, "rec_train: null tensor");
  PLAS_REQUIRE(backward ? (d->dout != nullptr && d->c_save != nullptr) : d->out != nullptr, "rec_train: null tensor");
  PLAS_REQUIRE(d->U % 4 == 0 && d->U <= 1024, "rec_train: U=%d unsupported (multiple of 4, <= 1024)", d->U);
  PLAS_REQUIRE(d->out_batch_stride >= (int64_t)d->T * d->ndir * d->U, "rec_train: out_batch_stride too small");
  size_t o_ctr, o_x, total;
  rt_ws_layout(*d, &o_ctr, &o_x, &total);
  PLAS_REQUIRE(workspace_bytes >= total, "rec_train: workspace %zu < %zu", workspace_bytes, total);

  // preferred: one portable-size cluster (<= 8 CTAs) per (direction, group), exchange through distributed shared memory
  const char* mode = getenv("PLAS_RT_EXCHANGE");  // "l2" forces the counter protocol (tuning aid)
  if (!backward && d->U % RT_CL_UPC == 0 && d->U / RT_CL_UPC <= 8 && d->U % (4 * RT_CL_KS) == 0 && !(mode && mode[0] == 'l')) {
    const int G = d->U / RT_CL_UPC;
    const size_t smem = (size_t)d->U * RT_CL_UPC * 16 + (size_t)2 * RT_CL_R * (d->U + 4) * 4 +
                        (size_t)RT_CL_KS * RT_CL_R * RT_CL_UPC * 16 + (size_t)RT_CL_R * RT_CL_UPC * 4;
    if (smem <= 220 * 1024) {
      const void* fn = (const void*)rec_train_fwd_kernel<RT_CL_R, RT_CL_UPC, RT_CL_RB, true>;
      PLAS_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      RecTrainArgs a;
      a.d = *d;
      a.counters = nullptr;
      a.xbuf = nullptr;
      a.n_groups = (d->B + RT_CL_R - 1) / RT_CL_R;
      a.G = G;
      a.Bpad = a.n_groups * RT_CL_R;
      a.group_offset = 0;
      a.groups_here = a.n_groups;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)(d->ndir * a.n_groups * G));
      cfg.blockDim = dim3(RT_THREADS);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = (unsigned)G;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      // only when every cluster is co-resident (one wave): with more groups than that the L2 variants, which spread a
      // group over more CTAs, finish sooner (measured at B = 64: 5.1 ms vs 3.5 ms)
      int max_clusters = 0;
      if (cudaOccupancyMaxActiveClusters(&max_clusters, fn, &cfg) != cudaSuccess) {
        (void)cudaGetLastError();
        max_clusters = 0;
      }
      if (d->ndir * a.n_groups <= max_clusters) {
        PLAS_CUDA(cudaLaunchKernelEx(&cfg, rec_train_fwd_kernel<RT_CL_R, RT_CL_UPC, RT_CL_RB, true>, a));
        return PLAS_OK;
      }
    }
  }
  // otherwise L2 exchange: pick the shape with the fewest launches (all CTAs of a group must be co-resident), then the most CTAs
  const RtShape* shapes = backward ? rt_bwd_shapes : rt_fwd_shapes;
  const int n_shapes = 6;
  const RtShape* best = nullptr;
  int best_launches = 1 << 30, best_gpl = 0;
  size_t best_smem = 0;
  for (int i = 0; i < n_shapes; ++i) {
    const RtShape& sh = shapes[i];
    if (!rt_shape_ok(sh, d->U, backward)) continue;
    const size_t smem = rt_smem_bytes(sh, d->U, backward);
    PLAS_CUDA(cudaFuncSetAttribute(sh.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    PLAS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sh.fn, RT_THREADS, smem));
    const int resident = per_sm * num_sms();
    const int ctas_per_group = d->ndir * (d->U / sh.UPC);
    const int gpl = resident / ctas_per_group;
    if (gpl < 1) continue;
    const int n_groups = (d->B + sh.R - 1) / sh.R;
    const int launches = (n_groups + gpl - 1) / gpl;
    if (launches < best_launches) {
      best = &sh; best_launches = launches; best_gpl = gpl; best_smem = smem;
    }
  }
  PLAS_REQUIRE(best != nullptr, "rec_train: no kernel shape fits U=%d", d->U);
  RecTrainArgs a;
  a.d = *d;
  a.counters = (unsigned*)((unsigned char*)workspace + o_ctr);
  a.xbuf = (float*)((unsigned char*)workspace + o_x);
  a.n_groups = (d->B + best->R - 1) / best->R;
  a.G = d->U / best->UPC;
  a.Bpad = a.n_groups * best->R;
  PLAS_CUDA(cudaMemsetAsync(a.counters, 0, (size_t)d->ndir * a.n_groups * 4, stream));
  for (int g0 = 0; g0 < a.n_groups; g0 += best_gpl) {
    a.group_offset = g0;
    a.groups_here = (a.n_groups - g0 < best_gpl) ? (a.n_groups - g0) : best_gpl;
    void* args[] = {&a};
    PLAS_CUDA(cudaLaunchCooperativeKernel(best->fn, dim3(d->ndir * a.G * a.groups_here), dim3(RT_THREADS), args, best_smem, stream));
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
