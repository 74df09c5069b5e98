// K4 -- attention decoder, every decode step on the device (persistent cooperative kernel).
//
// Replaces AttentionWrapper(MultiRNNCell([LSTMCell]*L), {Luong,Bahdanau,LuongMonotonic}Attention)
// + BasicDecoder + GreedyEmbeddingHelper / TrainingHelper + dynamic_decode (reference
// las/model.py:145-202, 205-349) and the DenseBinfDecoder projection
// (utils/training_helper.py:122-153).  Wiring = the reference defaults: bottom_only=False,
// attention_layer_size=None (the wrapper emits the context vector), one-hot inputs.
//
// One step = L+1 phases separated by a grid barrier (L2 atomics, acquire/release):
//   phase A(l): LSTM layer l for all rows.  A CTA owns 4 hidden units (16 gate columns); its 8
//               warps split K, partial sums meet in shared memory and the gate math is fused
//               (bf16: mma.m16n8k16 with pre-packed B fragments streamed from L2; f32: FFMA).
//   phase B   : one CTA per utterance: query layer, scores, masked softmax / monotonic
//               recurrence, context reduction over the encoder memory, output projection,
//               greedy argmax (lowest index wins ties) and the finished / sequence-length logic.
// Step-to-step activations live in small L2-resident ping-pong buffers read with ld.global.cg.
#include "common.cuh"
#include "../../include/plas.h"

namespace plas {

constexpr int DEC_THREADS = 256;
constexpr int DEC_ROWS = 64;  // rows per phase-A pass

struct DecArgs {
  plas_dec_desc d;
  unsigned char* xbuf[4];   // per layer: [2][B][K_l] activations (dtype)
  float* c_state;           // [L][B][Ud]
  float* align_state;       // [B][Tm] (monotonic)
  int* cur_ids;             // [B]
  int* finished;            // [B]
  unsigned* n_finished;     // [1]
  unsigned* bar;            // [1] grid barrier counter
};

__device__ __forceinline__ float tanh_fast(float x) {  // 1 - 2/(e^{2x}+1), abs err ~1e-7
  return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f);
}

template <typename T> __device__ __forceinline__ float ldcg_f(const T* p);
template <> __device__ __forceinline__ float ldcg_f<float>(const float* p) { return __ldcg(p); }
template <> __device__ __forceinline__ float ldcg_f<__nv_bfloat16>(const __nv_bfloat16* p) {
  const unsigned short u = __ldcg(reinterpret_cast<const unsigned short*>(p));
  return __uint_as_float((unsigned)u << 16);
}

__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += 1;
    __threadfence();
    red_release_add_u32(bar, 1u);
    const unsigned target = epoch * gridDim.x;
    unsigned spins = 0;
    while (ld_acquire_u32(bar) < target) {
      if (++spins > (1u << 28)) __trap();
    }
    __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ void mma16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// 16 x 16 gate columns partial products of one slice into s_part[warp][row][16]
template <typename AT>
__device__ __forceinline__ void cell_partials(const AT* __restrict__ X, int K, int row_base, int nrows,
                                              const void* __restrict__ wslice, float* s_part, int warp, int lane);

template <>
__device__ __forceinline__ void cell_partials<__nv_bfloat16>(const __nv_bfloat16* __restrict__ X, int K,
                                                             int row_base, int nrows,
                                                             const void* __restrict__ wslice, float* s_part,
                                                             int warp, int lane) {
  const int g = lane >> 2, q = lane & 3;
  const int KS = K / 16;
  const int n_mt = (nrows + 15) / 16;
  float acc[4][2][4];
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[m][n][i] = 0.f;
  const uint4* wf = reinterpret_cast<const uint4*>(wslice) + lane;
  for (int ks = warp; ks < KS; ks += 8) {
    const uint4 w = __ldg(wf + (size_t)ks * 32);
    const int k0 = ks * 16 + 2 * q;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      if (m < n_mt) {
        const int r0 = m * 16 + g, r1 = r0 + 8;
        uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        if (r0 < nrows) {
          const unsigned* p = reinterpret_cast<const unsigned*>(X + (size_t)(row_base + r0) * K + k0);
          a0 = __ldcg(p);
          a2 = __ldcg(p + 4);
        }
        if (r1 < nrows) {
          const unsigned* p = reinterpret_cast<const unsigned*>(X + (size_t)(row_base + r1) * K + k0);
          a1 = __ldcg(p);
          a3 = __ldcg(p + 4);
        }
        mma16816(acc[m][0], a0, a1, a2, a3, w.x, w.y);
        mma16816(acc[m][1], a0, a1, a2, a3, w.z, w.w);
      }
    }
  }
  float* part = s_part + (size_t)warp * DEC_ROWS * 16;
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    if (m < n_mt) {
      const int r0 = m * 16 + g, r1 = r0 + 8;
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        part[r0 * 16 + n * 8 + 2 * q] = acc[m][n][0];
        part[r0 * 16 + n * 8 + 2 * q + 1] = acc[m][n][1];
        part[r1 * 16 + n * 8 + 2 * q] = acc[m][n][2];
        part[r1 * 16 + n * 8 + 2 * q + 1] = acc[m][n][3];
      }
    }
  }
}

template <>
__device__ __forceinline__ void cell_partials<float>(const float* __restrict__ X, int K, int row_base,
                                                     int nrows, const void* __restrict__ wslice, float* s_part,
                                                     int warp, int lane) {
  // f32 layout: wslice[k][16], column = 4*unit_local + gate.  lane -> (column, row parity)
  const int c = lane & 15, rh = lane >> 4;
  const float* W = reinterpret_cast<const float*>(wslice);
  float acc[DEC_ROWS / 2];
#pragma unroll
  for (int i = 0; i < DEC_ROWS / 2; ++i) acc[i] = 0.f;
  const int kper = (K + 7) / 8;
  const int kbeg = warp * kper, kend = min(K, kbeg + kper);
  for (int k = kbeg; k < kend; ++k) {
    const float w = __ldg(W + (size_t)k * 16 + c);
#pragma unroll
    for (int i = 0; i < DEC_ROWS / 2; ++i) {
      const int r = rh + 2 * i;
      if (r < nrows) acc[i] = fmaf(__ldcg(X + (size_t)(row_base + r) * K + k), w, acc[i]);
    }
  }
  float* part = s_part + (size_t)warp * DEC_ROWS * 16;
#pragma unroll
  for (int i = 0; i < DEC_ROWS / 2; ++i) {
    const int r = rh + 2 * i;
    if (r < nrows) part[r * 16 + c] = acc[i];
  }
}

// gate-column position inside a slice's 16 partial columns
template <typename AT> __device__ __forceinline__ int gate_col(int ul, int gate);
template <> __device__ __forceinline__ int gate_col<__nv_bfloat16>(int ul, int gate) {
  // n-tile A holds (i,j) pairs, n-tile B holds (f,o) pairs: [i0 j0 i1 j1 i2 j2 i3 j3 | f0 o0 ...]
  return (gate >> 1) * 8 + 2 * ul + (gate & 1);
}
template <> __device__ __forceinline__ int gate_col<float>(int ul, int gate) { return 4 * ul + gate; }

template <typename AT>
__global__ void __launch_bounds__(DEC_THREADS, 1) decoder_kernel(DecArgs p) {
  extern __shared__ __align__(16) unsigned char dec_smem[];
  const plas_dec_desc& d = p.d;
  const int B = d.B, Tm = d.Tm, D = d.D, Ud = d.Ud, V = d.V, L = d.n_layers;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr bool kBf16 = sizeof(AT) == 2;
  constexpr size_t wbytes_per_k16 = kBf16 ? 32 * 16 : 16 * 16 * 4;  // bytes of one slice per 16 k

  // shared memory: phase A partials [8][64][16] f32 (32 KB); phase B reuses the same bytes
  float* s_part = reinterpret_cast<float*>(dec_smem);
  float* s_hq = reinterpret_cast<float*>(dec_smem);  // [Ud]
  float* s_q = s_hq + Ud;                            // [Ud]
  float* s_score = s_q + Ud;                         // [Tm]
  float* s_att = s_score + Tm;                       // [D]
  float* s_logit = s_att + D;                        // [V]
  float* s_red = s_logit + V;                        // [32]

  // greedy stop: rint(max(mem_len) * factor)  (las/model.py:270-274)
  int max_iter = d.max_steps;
  if (!d.teacher_forced) {
    int ml = 0;
    for (int b = 0; b < B; ++b) ml = max(ml, d.mem_len[b]);
    const int mi = (int)rintf((float)ml * d.decoding_length_factor);
    max_iter = min(max_iter, mi);
  }
  unsigned epoch = 0;
  int t = 0;
  for (; t < max_iter; ++t) {
    if (!d.teacher_forced && __ldcg(p.n_finished) >= (unsigned)B) break;
    const int par = t & 1;
    // ---------------- phase A: LSTM layers ----------------
    for (int l = 0; l < L; ++l) {
      const int K = (l == 0) ? (D + Ud) : 2 * Ud;
      const AT* X = reinterpret_cast<const AT*>(p.xbuf[l]) + (size_t)par * B * K;
      AT* Xn_self = reinterpret_cast<AT*>(p.xbuf[l]) + (size_t)(par ^ 1) * B * K + (K - Ud);
      AT* Xn_up = nullptr;
      int Kup = 0;
      if (l + 1 < L) {
        Kup = 2 * Ud;
        Xn_up = reinterpret_cast<AT*>(p.xbuf[l + 1]) + (size_t)par * B * Kup;
      }
      const int nsl = Ud / 4;
      for (int slice = blockIdx.x; slice < nsl; slice += gridDim.x) {
        const unsigned char* wslice =
            reinterpret_cast<const unsigned char*>(d.w_cell[l]) + (size_t)slice * (K / 16) * wbytes_per_k16;
        for (int rb = 0; rb < B; rb += DEC_ROWS) {
          const int nrows = min(DEC_ROWS, B - rb);
          cell_partials<AT>(X, K, rb, nrows, wslice, s_part, warp, lane);
          __syncthreads();
          {
            const int r = tid >> 2, ul = tid & 3;
            if (r < nrows) {
              const int b = rb + r;
              const int u = slice * 4 + ul;
              float z[4];
#pragma unroll
              for (int gt = 0; gt < 4; ++gt) {
                float s = d.b_cell[l][4 * u + gt];
                const int col = gate_col<AT>(ul, gt);
#pragma unroll
                for (int w = 0; w < 8; ++w) s += s_part[((size_t)w * DEC_ROWS + r) * 16 + col];
                z[gt] = s;
              }
              if (l == 0) {
                int id;
                if (d.teacher_forced) id = d.forced_ids[(size_t)b * d.max_steps + t];
                else id = (t == 0) ? d.sos_id : __ldcg(p.cur_ids + b);
                id = max(0, min(id, V - 1));
                const AT* er = reinterpret_cast<const AT*>(d.w_emb) + (size_t)id * 4 * Ud + 4 * u;
#pragma unroll
                for (int gt = 0; gt < 4; ++gt) z[gt] += to_f32<AT>(er[gt]);
              }
              float* cp = p.c_state + ((size_t)l * B + b) * Ud + u;
              float cn, hn;
              lstm_gates(z[0], z[1], z[2], z[3], __ldcg(cp), cn, hn);
              __stcg(cp, cn);
              const AT hq = from_f32<AT>(hn);
              Xn_self[(size_t)b * K + u] = hq;
              if (Xn_up) Xn_up[(size_t)b * Kup + u] = hq;
            }
          }
          __syncthreads();
        }
      }
      grid_barrier(p.bar, epoch);
    }
    // ---------------- phase B: attention + projection + sampling ----------------
    {
      const int Ktop = (L == 1) ? (D + Ud) : 2 * Ud;
      const AT* Htop = reinterpret_cast<const AT*>(p.xbuf[L - 1]) + (size_t)(par ^ 1) * B * Ktop + (Ktop - Ud);
      for (int b = blockIdx.x; b < B; b += gridDim.x) {
        const int len = min(d.mem_len[b], Tm);
        for (int u = tid; u < Ud; u += DEC_THREADS) s_hq[u] = ldcg_f<AT>(Htop + (size_t)b * Ktop + u);
        __syncthreads();
        const AT* keys = reinterpret_cast<const AT*>(d.keys) + (size_t)b * Tm * Ud;
        if (d.attention_type == PLAS_ATT_BAHDANAU) {
          const AT* wq = reinterpret_cast<const AT*>(d.w_query);
          for (int u = tid; u < Ud; u += DEC_THREADS) {
            float acc = 0.f;
            for (int k = 0; k < Ud; ++k) acc = fmaf(s_hq[k], to_f32<AT>(wq[(size_t)k * Ud + u]), acc);
            s_q[u] = acc;
          }
          __syncthreads();
        }
        // scores: warp per memory position
        for (int tm = warp; tm < Tm; tm += 8) {
          float sc = -INFINITY;
          if (tm < len) {
            const AT* kr = keys + (size_t)tm * Ud;
            float acc = 0.f;
            if (d.attention_type == PLAS_ATT_BAHDANAU) {
              for (int u = lane; u < Ud; u += 32) acc = fmaf(d.v_att[u], tanh_fast(to_f32<AT>(kr[u]) + s_q[u]), acc);
            } else {
              for (int u = lane; u < Ud; u += 32) acc = fmaf(to_f32<AT>(kr[u]), s_hq[u], acc);
            }
            sc = warp_sum(acc);
          }
          if (lane == 0) s_score[tm] = sc;
        }
        __syncthreads();
        if (d.attention_type == PLAS_ATT_LUONG_MONOTONIC) {
          if (tid == 0) {
            // tf.contrib.seq2seq.monotonic_attention(mode='parallel') evaluated sequentially
            float* prev = p.align_state + (size_t)b * Tm;
            const float tiny = 1.17549435e-38f;
            float cs = 0.f, csum = 0.f;
            for (int tm = 0; tm < Tm; ++tm) {
              const float pc = (tm < len) ? sigmoidf_acc(s_score[tm] + d.score_bias) : 0.f;
              const float cpv = expf(cs);
              const float pv = (t == 0) ? (tm == 0 ? 1.f : 0.f) : __ldcg(prev + tm);
              csum += pv / fminf(fmaxf(cpv, 1e-10f), 1.f);
              const float a = pc * cpv * csum;
              cs += logf(fminf(fmaxf(1.f - pc, tiny), 1.f));
              s_score[tm] = a;
            }
          }
          __syncthreads();
          for (int tm = tid; tm < Tm; tm += DEC_THREADS) __stcg(p.align_state + (size_t)b * Tm + tm, s_score[tm]);
        } else {
          float m = -INFINITY;
          for (int tm = tid; tm < Tm; tm += DEC_THREADS) m = fmaxf(m, s_score[tm]);
          m = warp_max(m);
          if (lane == 0) s_red[warp] = m;
          __syncthreads();
          m = s_red[0];
          for (int w = 1; w < 8; ++w) m = fmaxf(m, s_red[w]);
          __syncthreads();
          float sum = 0.f;
          for (int tm = tid; tm < Tm; tm += DEC_THREADS) {
            const float e = (tm < len) ? expf(s_score[tm] - m) : 0.f;
            s_score[tm] = e;
            sum += e;
          }
          sum = warp_sum(sum);
          if (lane == 0) s_red[warp] = sum;
          __syncthreads();
          sum = 0.f;
          for (int w = 0; w < 8; ++w) sum += s_red[w];
          for (int tm = tid; tm < Tm; tm += DEC_THREADS) s_score[tm] = s_score[tm] / sum;
          __syncthreads();
        }
        if (d.alignment) {
          float* ar = d.alignment + ((size_t)b * d.max_steps + t) * Tm;
          for (int tm = tid; tm < Tm; tm += DEC_THREADS) ar[tm] = s_score[tm];
        }
        // context = sum_t a[t] * values[b,t,:]   (thread owns 8 consecutive channels)
        const AT* vals = reinterpret_cast<const AT*>(d.values) + (size_t)b * Tm * D;
        AT* att_out = reinterpret_cast<AT*>(p.xbuf[0]) + (size_t)(par ^ 1) * B * (D + Ud) + (size_t)b * (D + Ud);
        for (int ch = tid; ch < D / 8; ch += DEC_THREADS) {
          float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          for (int tm = 0; tm < len; ++tm) {
            const float a = s_score[tm];
            const AT* vp = vals + (size_t)tm * D + ch * 8;
            if constexpr (kBf16) {
              const uint4 raw = __ldg(reinterpret_cast<const uint4*>(vp));
              const unsigned w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                acc[2 * i] = fmaf(a, __uint_as_float(w[i] << 16), acc[2 * i]);
                acc[2 * i + 1] = fmaf(a, __uint_as_float(w[i] & 0xffff0000u), acc[2 * i + 1]);
              }
            } else {
              const float4 v0 = __ldg(reinterpret_cast<const float4*>(vp));
              const float4 v1 = __ldg(reinterpret_cast<const float4*>(vp) + 1);
              acc[0] = fmaf(a, v0.x, acc[0]); acc[1] = fmaf(a, v0.y, acc[1]);
              acc[2] = fmaf(a, v0.z, acc[2]); acc[3] = fmaf(a, v0.w, acc[3]);
              acc[4] = fmaf(a, v1.x, acc[4]); acc[5] = fmaf(a, v1.y, acc[5]);
              acc[6] = fmaf(a, v1.z, acc[6]); acc[7] = fmaf(a, v1.w, acc[7]);
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const AT qv = from_f32<AT>(acc[i]);
            att_out[ch * 8 + i] = qv;   // fed back to the cell in the storage dtype ...
            s_att[ch * 8 + i] = acc[i];  // ... the projection consumes the f32 context
          }
        }
        __syncthreads();
        // logits[v] = att . WpT[v] + b   (warp per output, WpT is [V][D])
        const AT* wpt = reinterpret_cast<const AT*>(d.w_proj);
        for (int v = warp; v < V; v += 8) {
          const AT* wr = wpt + (size_t)v * D;
          float acc = 0.f;
          for (int k = lane; k < D; k += 32) acc = fmaf(s_att[k], to_f32<AT>(wr[k]), acc);
          acc = warp_sum(acc);
          if (lane == 0) s_logit[v] = acc + d.b_proj[v];
        }
        __syncthreads();
        float* lrow = d.logits + ((size_t)b * d.max_steps + t) * V;
        for (int v = tid; v < V; v += DEC_THREADS) lrow[v] = s_logit[v];
        if (warp == 0) {
          float best = -INFINITY;
          int bi = 0x7fffffff;
          for (int v = lane; v < V; v += 32) {
            const float x = s_logit[v];
            if (x > best) { best = x; bi = v; }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
          }
          if (lane == 0) {
            if (bi == 0x7fffffff) bi = 0;
            d.sample_ids[(size_t)b * d.max_steps + t] = bi;
            if (!d.teacher_forced) {
              const int was = __ldcg(p.finished + b);
              if (!was) d.seq_len[b] = t + 1;
              const int now = was || (bi == d.eos_id) || (t + 1 >= max_iter);
              if (now && !was) {
                __stcg(p.finished + b, 1);
                atomicAdd(p.n_finished, 1u);
              }
              __stcg(p.cur_ids + b, bi);
            }
          }
        }
        __syncthreads();
      }
      grid_barrier(p.bar, epoch);
    }
  }
  if (blockIdx.x == 0 && tid == 0) *d.n_steps = t;
}

static void dec_ws_layout(const plas_dec_desc& d, size_t* offs, size_t* total) {
  // offs: 0..3 xbuf[l], 4 c_state, 5 align_state, 6 cur_ids, 7 finished, 8 n_finished, 9 bar
  const size_t esz = d.dtype == PLAS_BF16 ? 2 : 4;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~size_t(255); return o; };
  for (int l = 0; l < 4; ++l) {
    const size_t K = (l == 0) ? (size_t)d.D + d.Ud : (size_t)2 * d.Ud;
    offs[l] = take(l < d.n_layers ? 2 * (size_t)d.B * K * esz : 0);
  }
  offs[4] = take((size_t)d.n_layers * d.B * d.Ud * 4);
  offs[5] = take((size_t)d.B * d.Tm * 4);
  offs[6] = take((size_t)d.B * 4);
  offs[7] = take((size_t)d.B * 4);
  offs[8] = take(4);
  offs[9] = take(4);
  *total = off;
}

size_t dec_tc_workspace_bytes(const plas_dec_desc& d);
int dec_tc_launch(const plas_dec_desc& d, void* workspace, size_t workspace_bytes, cudaStream_t stream);
size_t dec_fold_workspace_bytes(const plas_dec_desc& d);
int dec_fold_launch(const plas_dec_desc& d, void* workspace, size_t workspace_bytes, cudaStream_t stream);

}  // namespace plas

using namespace plas;

extern "C" size_t plas_decoder_workspace_bytes(const plas_dec_desc* d) {
  size_t offs[10], total;
  dec_ws_layout(*d, offs, &total);
  const size_t tc = dec_tc_workspace_bytes(*d);
  const size_t fold = dec_fold_workspace_bytes(*d);
  if (tc > total) total = tc;
  return fold > total ? fold : total;
}

extern "C" int plas_decoder_fwd(const plas_dec_desc* d, void* workspace, size_t workspace_bytes,
                                plas_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PLAS_REQUIRE(d && workspace, "decoder: null argument");
  PLAS_REQUIRE(d->dtype == PLAS_F32 || d->dtype == PLAS_BF16, "decoder: dtype %d", d->dtype);
  PLAS_REQUIRE(d->B > 0 && d->Tm > 0 && d->V > 0 && d->max_steps >= 0, "decoder: bad shape");
  PLAS_REQUIRE(d->n_layers >= 1 && d->n_layers <= 4, "decoder: n_layers=%d (1..4)", d->n_layers);
  PLAS_REQUIRE(d->Ud % 16 == 0 && d->D % 16 == 0, "decoder: Ud=%d and D=%d must be multiples of 16", d->Ud, d->D);
  PLAS_REQUIRE((d->attention_type >= 0 && d->attention_type <= 2) || d->attention_type == PLAS_ATT_CUSTOM, "decoder: attention_type %d",
               d->attention_type);
  PLAS_REQUIRE(d->keys && d->values && d->mem_len && d->w_emb && d->w_proj && d->b_proj && d->logits &&
                   d->sample_ids && d->seq_len && d->n_steps, "decoder: null tensor");
  PLAS_REQUIRE(d->attention_type != PLAS_ATT_BAHDANAU || (d->w_query && d->v_att), "decoder: bahdanau needs query_layer/attention_v");
  PLAS_REQUIRE(!d->teacher_forced || d->forced_ids, "decoder: teacher forcing needs forced_ids");
  for (int l = 0; l < d->n_layers; ++l) PLAS_REQUIRE(d->w_cell[l] && d->b_cell[l], "decoder: layer %d weights missing", l);
  {
    int rc = dec_fold_launch(*d, workspace, workspace_bytes, stream);
    if (rc != 1) return rc;  // 1 = shape not eligible for the folded-context tensor-core kernel
    if (d->attention_type == PLAS_ATT_CUSTOM)  // only decoder_fold.cu carries it here (any shape: plas_decoder_infer_f32)
      return set_err(PLAS_EUNSUPPORTED, "decoder: custom attention runs on the folded tensor-core kernel only (bf16, B <= 128, "
                     "D and Ud multiples of 64, w_query_tc / vw / pv supplied); use plas_decoder_infer_f32 otherwise");
    rc = dec_tc_launch(*d, workspace, workspace_bytes, stream);
    if (rc != 1) return rc;  // 1 = shape not eligible for the tensor-core kernel
  }
  size_t offs[10], total;
  dec_ws_layout(*d, offs, &total);
  PLAS_REQUIRE(workspace_bytes >= total, "decoder: workspace %zu < %zu", workspace_bytes, total);
  PLAS_CUDA(cudaMemsetAsync(workspace, 0, total, stream));
  PLAS_CUDA(cudaMemsetAsync(d->seq_len, 0, (size_t)d->B * 4, stream));
  PLAS_CUDA(cudaMemsetAsync(d->n_steps, 0, 4, stream));
  if (d->max_steps == 0) return PLAS_OK;

  DecArgs a;
  a.d = *d;
  unsigned char* ws = (unsigned char*)workspace;
  for (int l = 0; l < 4; ++l) a.xbuf[l] = ws + offs[l];
  a.c_state = (float*)(ws + offs[4]);
  a.align_state = (float*)(ws + offs[5]);
  a.cur_ids = (int*)(ws + offs[6]);
  a.finished = (int*)(ws + offs[7]);
  a.n_finished = (unsigned*)(ws + offs[8]);
  a.bar = (unsigned*)(ws + offs[9]);

  size_t smem_b = (size_t)(2 * d->Ud + d->Tm + d->D + d->V + 32) * 4;
  size_t smem = (size_t)8 * DEC_ROWS * 16 * 4;
  if (smem_b > smem) smem = smem_b;
  PLAS_REQUIRE(smem <= 200 * 1024, "decoder: needs %zu bytes of shared memory", smem);
  const void* fn = d->dtype == PLAS_BF16 ? (const void*)decoder_kernel<__nv_bfloat16> : (const void*)decoder_kernel<float>;
  PLAS_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  PLAS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, DEC_THREADS, smem));
  PLAS_REQUIRE(per_sm >= 1, "decoder: kernel does not fit on an SM");
  int grid = num_sms();
  const int want = (d->Ud / 4 > d->B) ? d->Ud / 4 : d->B;
  if (grid > want) grid = want;
  void* args[] = {&a};
  PLAS_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(DEC_THREADS), args, smem, stream));
  return PLAS_OK;
}
