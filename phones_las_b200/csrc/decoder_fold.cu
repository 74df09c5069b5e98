// K4 (tensor-core path, bf16, folded context) -- weight-stationary attention decoder for sm_100a.
//
// Same contract as decoder.cu / decoder_tc.cu (AttentionWrapper(MultiRNNCell) + BasicDecoder + Greedy/Training helper +
// dynamic_decode, reference las/model.py:145-349; projection utils/training_helper.py:122-153) for the default wiring,
// reorganised so that a decode step costs two or three grid barriers instead of four and streams no explicit context:
//
//  * The context never exists.  AttentionWrapper feeds attention_{t-1} = sum_t a_t values[t] back into cell 0
//    (las/model.py:195-200), i.e. cell 0 adds context . W0[V:V+D].  That product is linear in the values, so
//    VW = values . W0[V:V+D]  ([B][Tm][4Ud], bf16, one K2 GEMM per batch -- like the existing PV = values . W_proj)
//    lets the attention phase emit cell 0's pre-activation part  zctx = sum_t a_t VW[t]  directly.  VW has the width
//    of the gate vector (4Ud) instead of D; cell 0's per-step contraction shrinks from K = D + Ud to K = Ud.
//  * Cell 0 moves into the attention phase.  The CTAs that own an utterance's attention also own its cell-0 state:
//    right after the logits/argmax of step t they add zctx + W_emb[id_t] + (h0_t . W0h, computed one phase earlier)
//    + b0, apply the gates and publish h0_{t+1}.  The former "LSTM-0" phase and its grid barrier disappear.
//  * Every remaining GEMM phase has A = h_l (K = Ud): the phase that streams h_l multiplies it with everything that
//    reads it -- layer l+1's input rows (consumed now), layer l's own recurrent rows (carried in registers to the
//    next step; for l = 0 written to ZH0 for the attention CTAs) and, for the top layer, the bahdanau query layer.
//    Step = [GEMM(h_0)] B ... [GEMM(h_{L-1})] B [attention + cell 0] B; the barrier after the top phase is only
//    needed when something reads its result in the same step (bahdanau query; L = 1).
//  * Attention work item = (utterance, part): AS = 2 or 4 CTAs of one cluster share an utterance.  They split the
//    depth of the score reduction (partial scores swapped through DSMEM), the 4Ud columns of zctx (= whole hidden
//    units of cell 0) and the memory rows of the logits  a . PV  (partial logits swapped through DSMEM, so every
//    CTA of the item knows the argmax without a trip through global memory).  When each CTA has a single item its
//    slice of the keys stays resident in shared memory for the whole decode (c2: 96 KB per CTA).
//  * GEMM phases as in decoder_tc.cu: weight slices resident in shared memory as K-major SWIZZLE_128B UMMA B tiles,
//    activations by TMA, tcgen05.mma M=128 with the accumulator in TMEM, epilogue thread r = batch row r.  Clusters
//    of 4 CTAs split K four ways when Ud % 256 == 0 and exchange partial sums through DSMEM (fixed order).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tcgen05.cuh"
#include "../../include/plas.h"

namespace plas {

constexpr int DF_THREADS = 256;   // 8 warps (2 per scheduler: 255 registers each); GEMM phases: warp 0 MMA issuer, warp 1 copy
                                  // producer, warps 4..7 epilogue rows; attention phase: all 8 warps
constexpr int DF_MAX_STAGES = 16;
constexpr int DF_CL = 4;          // cluster size (always)

struct alignas(64) DecFoldArgs {
  plas_dec_desc d;
  const unsigned char* w_x[4];   // l >= 1: [Ud/4][Ud/64][2048 B] swizzled UMMA B tiles of the input rows of layer l
  const unsigned char* w_h[4];   // recurrent rows of layer l, same layout
  const unsigned char* w_q;      // [Ud/16][Ud/64][2048 B] bahdanau query layer
  unsigned char* hbuf[4];        // [2 parities][Ud/64 k blocks][RT rows][64] bf16: K-major SWIZZLE_128B UMMA A tiles, written in
                                 // place by the producers of h so that a k block is ONE contiguous bulk copy
  float* qbuf;                   // [B][Ud]
  float* zh0;                    // [B][4Ud]  h0_t . W0h (unit-major columns)
  float* c0;                     // [B][Ud]   cell-0 state
  float* align_state;            // [2][B][Tm] monotonic alignments by step parity (step t reads t-1's, writes its own)
  int* next_ids;                 // [B] argmax of the step just decoded
  unsigned* bar;
  unsigned long long* dbg;       // optional phase timers (ns summed over steps), CTA 0
  int as;                        // CTAs per utterance in the attention phase (2 or 4)
  int m64;                       // B <= 64: MMAs with M = 64
  int keys_res;                  // this CTA's key slice is resident in shared memory
  int n_stages_a, stage_a;       // activation ring of the GEMM phases; stage_a = tile bytes = RT * 128
  int off_w[4];                  // per phase: resident weight tiles
  int off_keys, off_ring, off_comb, off_pv, pv_cap, off_att, off_misc;
  int tm_pad, v_pad;
};

namespace {

__device__ __forceinline__ uint32_t df_mapa(uint32_t local_saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void df_st_async_v4(uint32_t raddr, const uint4 v, uint32_t rbar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(rbar)
               : "memory");
}
__device__ __forceinline__ void df_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void df_fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void df_sync() { __syncthreads(); }

// Grid barrier between phases.  What crosses it into the ASYNC proxy: h tiles written with generic stores to global memory and
// then read by the next phase's bulk copies, and the shared-memory region the phases share (generic accesses of one phase, bulk
// copy writes of the next).  Writer side: every thread orders its generic global writes against the async proxy
// (fence.proxy.async.global) before the release; reader side: only the warp that issues the bulk copies needs the full proxy
// fence after the acquire.  (A full `fence.proxy.async` by all 256 threads on both sides was 20 % of the kernel's stall samples.)
__device__ __forceinline__ void df_grid_barrier(unsigned* bar, unsigned& epoch) {
  asm volatile("fence.proxy.async.global;" ::: "memory");
  __syncthreads();         // every thread's writes happen-before thread 0's release (cumulative at gpu scope)
  if (threadIdx.x == 0) {
    epoch += 1;
    red_release_add_u32(bar, 1u);
    const unsigned target = epoch * gridDim.x;
    unsigned spins = 0;
    while (ld_acquire_u32(bar) < target) {
      if (++spins > (1u << 28)) __trap();
    }
  }
  __syncthreads();
  if ((threadIdx.x >> 5) == 1) df_fence_proxy_async();  // warp 1 = the copy producer of the GEMM phases
}

// bahdanau scores need B*Tm*Ud tanh per decode step (6.2 M at c2): one MUFU op each.  tanh.approx.f32 has a
// relative error of 2^-11, an eighth of the bf16 quantisation of the keys it is applied to.
__device__ __forceinline__ float df_tanh_mufu(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// byte offset of element (row r, unit u) inside an activation buffer of k-block tiles [Ud/64][RT][128 B] whose 16-byte chunks
// are XOR-swizzled by the row (what TMA's SWIZZLE_128B would write and the UMMA descriptor expects)
__device__ __forceinline__ size_t df_h_off(int r, int u, int tile_bytes) {
  return (size_t)(u >> 6) * tile_bytes + (size_t)r * 128 + (size_t)((((u & 63) >> 3) ^ (r & 7)) << 4) + (size_t)((u & 7) << 1);
}

// scores of 4 memory rows (r0, r0+8, r0+16, r0+24) of one warp from the rows' 16-byte key chunks: lane owns chunks lane (and
// lane+32 when NCC == 2) of this part's depth; q / v are zero where the lane has no chunk, so clamped loads are harmless
template <int NCC, bool BAHDANAU>
__device__ __forceinline__ void df_score4(const uint4 (*kk)[2], const float* qreg, const float* vreg, float* acc) {
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i] = 0.f;
#pragma unroll
  for (int cc = 0; cc < NCC; ++cc) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float k0[4], k1[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const unsigned kw = j == 0 ? kk[i][cc].x : (j == 1 ? kk[i][cc].y : (j == 2 ? kk[i][cc].z : kk[i][cc].w));
        k0[i] = __uint_as_float(kw << 16);
        k1[i] = __uint_as_float(kw & 0xffff0000u);
      }
      if (BAHDANAU) {
        // four rows with independent accumulators: the MUFU.TANH -> FFMA chains of different rows overlap
        float t0[4], t1[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          t0[i] = df_tanh_mufu(k0[i] + qreg[cc * 8 + 2 * j]);
          t1[i] = df_tanh_mufu(k1[i] + qreg[cc * 8 + 2 * j + 1]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[i] = fmaf(vreg[cc * 8 + 2 * j], t0[i], acc[i]);
          acc[i] = fmaf(vreg[cc * 8 + 2 * j + 1], t1[i], acc[i]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[i] = fmaf(k0[i], qreg[cc * 8 + 2 * j], acc[i]);
          acc[i] = fmaf(k1[i], qreg[cc * 8 + 2 * j + 1], acc[i]);
        }
      }
    }
  }
}

// all scores of one item: warp w takes rows w, w+8, ...; kb = this part's key slice (uint4 units, row stride kstride), in shared
// memory (RES) or global memory.  Loads are unconditional (row and chunk indices clamped into the slice): no divergent
// branches around them, the next group's loads are in flight while the current one is scored.
template <int NCC, bool BAHDANAU, bool RES>
__device__ __forceinline__ void df_scores(const uint4* kb, int kstride, int n_c8, int len, int warp, int lane,
                                          const float* qreg, const float* vreg, float* s_score) {
  int c8c[2];
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) c8c[cc] = min(lane + 32 * cc, n_c8 - 1);
  auto load4 = [&](int r0, uint4 (*kk)[2]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rc = min(r0 + 8 * i, len - 1);
#pragma unroll
      for (int cc = 0; cc < NCC; ++cc) {
        const uint4* src = kb + (size_t)rc * kstride + c8c[cc];
        kk[i][cc] = RES ? *src : __ldg(src);
      }
    }
  };
  auto finish4 = [&](int r0, float* acc) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + 8 * i;
      const float sacc = warp_sum(acc[i]);
      if (lane == 0 && r < len) s_score[r] = sacc;
    }
  };
  uint4 ka[4][2], kb2[4][2];
  float acc[4];
  load4(warp, ka);
  for (int r0 = warp; r0 < len; r0 += 64) {
    load4(r0 + 32, kb2);
    df_score4<NCC, BAHDANAU>(ka, qreg, vreg, acc);
    finish4(r0, acc);
    load4(r0 + 64, ka);
    if (r0 + 32 < len) {
      df_score4<NCC, BAHDANAU>(kb2, qreg, vreg, acc);
      finish4(r0 + 32, acc);
    }
  }
}

// In-place inclusive prefix sum of a[0..n) by the 256 threads of the CTA (fixed order: each thread sums a contiguous chunk, the
// chunk totals are scanned with warp shuffles, the eight warp totals through shared memory): three block barriers instead of
// the log2(n) of a Hillis-Steele scan.  The caller has synchronised after writing a[]; a[] is complete on return.
__device__ __forceinline__ void df_block_scan(float* a, int n, float* s_tot) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunk = (n + DF_THREADS - 1) / DF_THREADS;
  const int lo = min(tid * chunk, n), hi = min(lo + chunk, n);
  float tot = 0.f;
  for (int i = lo; i < hi; ++i) tot += a[i];
  float inc = tot;  // inclusive scan of the chunk totals inside the warp
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) s_tot[warp] = inc;
  __syncthreads();
  float off = inc - tot;  // exclusive prefix inside the warp
  for (int w = 0; w < warp; ++w) off += s_tot[w];
  for (int i = lo; i < hi; ++i) {
    off += a[i];
    a[i] = off;
  }
  __syncthreads();
}

}  // namespace

template <int L>
__global__ void __launch_bounds__(DF_THREADS, 1) decoder_fold_kernel(const __grid_constant__ DecFoldArgs p) {
  extern __shared__ unsigned char df_smem_raw[];
  const plas_dec_desc& d = p.d;
  const int B = d.B, Tm = d.Tm, Ud = d.Ud, V = d.V;
  const int W4 = 4 * Ud;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NSTA = p.n_stages_a, STA = p.stage_a;
  const uint32_t raw = smem_u32(df_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* smem = df_smem_raw + (base - raw);
  const uint32_t ring = base + (uint32_t)p.off_ring;

  // misc region: barriers, TMEM slot, small arrays
  unsigned char* misc = smem + p.off_misc;
  const uint32_t misc_u = base + (uint32_t)p.off_misc;
  auto fullA = [&](int s) { return misc_u + 8u * s; };
  const uint32_t wave_done = misc_u + 8u * DF_MAX_STAGES;  // the MMAs that read the current wave of ring stages completed
  const uint32_t tfull = wave_done + 8u;
  const uint32_t sbar = tfull + 8u;    // the partners' partial scores landed (tx bytes)
  const uint32_t lbar = tfull + 16u;   // the partners' partial logits landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 8 * (DF_MAX_STAGES + 6));
  float* s_bias = reinterpret_cast<float*>(misc + 256);          // [4][16] biases of layers 1..L-1 (this CTA's columns)
  float* s_v = s_bias + 64;                                      // [Ud] attention_v (lives for the whole decode)
  // scratch of the attention phase: inside the region the GEMM phases use as their activation ring (the phases alternate)
  float* s_comb = reinterpret_cast<float*>(smem + p.off_comb);   // [<= 2048] zctx partials of the row groups
  float* s_pv = reinterpret_cast<float*>(smem + p.off_pv);       // PV rows of this part (prefetched per step)
  float* s_red = reinterpret_cast<float*>(smem + p.off_att);     // [64]
  float* s_lp = s_red + 64;                                      // [4][64]
  float* s_lmine = s_lp + 256;                                   // [v_pad] my partial logits
  float* s_lpart = s_lmine + p.v_pad;                            // [4][v_pad] partial logits by part
  float* s_q = s_lpart + 4 * p.v_pad;                            // [Ud]
  float* s_score = s_q + Ud;                                     // [tm_pad]
  float* s_scan = s_score + p.tm_pad;                            // [2][tm_pad]
  float* s_peer = s_scan + 2 * p.tm_pad;                         // [4][tm_pad] partial scores by part

  const int nkb = Ud / 64;
  const int nq = Ud / 16;
  const bool bahdanau = d.attention_type == PLAS_ATT_BAHDANAU;
  const bool monotonic = d.attention_type == PLAS_ATT_LUONG_MONOTONIC;
  // CustomAttention (las/model.py:72-101): keys = relu(memory_layer(values)) (the caller passes them), query = relu(query_layer(h)),
  // luong score -- the query-layer machinery of the bahdanau path with the dot-product score of the luong path
  const bool custom = d.attention_type == PLAS_ATT_CUSTOM;
  const bool qlayer = bahdanau || custom;         // the top phase also multiplies h with a query layer
  const int slice = blockIdx.x;                   // this CTA owns gate columns 16*slice .. +15 of every layer
  const bool q_cta = qlayer && slice < nq;        // ... and query-layer columns 16*slice .. +15
  const int crank = blockIdx.x % DF_CL;           // rank inside the cluster
  const int AS = p.as;
  const int part = crank % AS;                    // which part of its utterances this CTA handles
  const int pbase = crank - part;                 // cluster rank of part 0 of my attention group
  const size_t hpar = (size_t)nkb * STA;          // bytes of one parity of an activation buffer

  // groups of a phase, in accumulator-column order: [x: input rows of layer ph+1] [h: recurrent rows of layer ph] [q]
  auto has_x = [&](int ph) { return ph < L - 1; };
  auto has_q = [&](int ph) { return ph == L - 1 && q_cta; };
  auto n_groups = [&](int ph) { return (has_x(ph) ? 1 : 0) + 1 + (has_q(ph) ? 1 : 0); };

  // ---- one-time setup: resident weight tiles, key slice, barriers, TMEM --------------------------------------
#pragma unroll
  for (int ph = 0; ph < L; ++ph) {
    const int ng = n_groups(ph);
    uint4* dst = reinterpret_cast<uint4*>(smem + p.off_w[ph]);
    // per k block the groups' 16-row tiles stack into one (16*ng)-row K-major UMMA B tile (8-row swizzle atoms, 1024 B apart)
    for (int i = tid; i < nkb * ng * 128; i += DF_THREADS) {
      const int w16 = i & 127;
      const int g = (i >> 7) % ng;
      const int kb = (i >> 7) / ng;
      const unsigned char* src;
      if (has_x(ph) && g == 0) src = p.w_x[ph + 1];
      else if (g == (has_x(ph) ? 1 : 0)) src = p.w_h[ph];
      else src = p.w_q;
      dst[i] = __ldg(reinterpret_cast<const uint4*>(src + ((size_t)slice * nkb + kb) * 2048) + w16);
    }
    if (ph >= 1 && tid < 16) s_bias[ph * 16 + tid] = d.b_cell[ph][slice * 16 + tid];
  }
  if (bahdanau)
    for (int u = tid; u < Ud; u += DF_THREADS) s_v[u] = d.v_att[u];
  const int Dk = Ud / AS;                          // key depth of a part
  const int n_c8 = Dk / 8;                         // its 16-byte chunks per memory row
  if (p.keys_res) {                                // single item per CTA: its key slice stays on chip
    const int b = blockIdx.x / AS;
    if (b < B) {
      const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(d.keys) + (size_t)b * Tm * Ud + part * Dk);
      uint4* dst = reinterpret_cast<uint4*>(smem + p.off_keys);
      for (int i = tid; i < Tm * n_c8; i += DF_THREADS) {
        const int r = i / n_c8, c = i - r * n_c8;
        dst[i] = __ldg(src + (size_t)r * (Ud / 8) + c);
      }
    }
  }
  if (tid == 0) {
    for (int s = 0; s < DF_MAX_STAGES; ++s) mbar_init(fullA(s), 1);
    mbar_init(wave_done, 1);
    mbar_init(tfull, 1);
    mbar_init(sbar, 1);
    mbar_init(lbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  df_fence_proxy_async();  // the resident weights were written with generic stores, UMMA reads them via the async proxy
  tc_fence_before();
  __syncthreads();
  // peers' mbarriers must be initialised before anyone signals them
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- replicated decode state ---------------------------------------------------------------------------------
  // epilogue threads (warps 4..7) own batch row `row` = their TMEM lane's row of D.  B <= 64 runs the MMAs with M = 64 (an
  // SS-mode MMA fetches its A rows from shared memory at about a row per clock whatever N is, so 64 rows cost half of 128):
  // D row i then sits in lane 32*(i/16) + i%16 -- the first 16 lanes of every warp quadrant (scripts/micro/m64_probe.cu)
  const bool m64 = p.m64 != 0;
  const bool row_thread = tid >= 128;
  const int row = m64 ? ((warp - 4) * 16 + lane) : (tid - 128);
  const bool row_valid = row_thread && row < B && (!m64 || lane < 16);
  float c_state[L][4];                            // cell states of layers 1..L-1 (index 0 unused)
  float carry[L][16];                             // h_{l,t-1} . W_lh for layers 1..L-1
#pragma unroll
  for (int l = 0; l < L; ++l) {
#pragma unroll
    for (int u = 0; u < 4; ++u) c_state[l][u] = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) carry[l][j] = 0.f;
  }
  int finished = 0;

  int max_iter = d.max_steps;
  if (!d.teacher_forced) {
    int ml = 0;
    for (int b = 0; b < B; ++b) ml = max(ml, d.mem_len[b]);
    max_iter = min(max_iter, (int)rintf((float)ml * d.decoding_length_factor));
  }

  uint32_t full_bits = 0, wave_parity = 0;  // mbarrier parities: ring stages (bit j, MMA issuer) / wave_done (producer)
  uint32_t acc_parity = 0;
  unsigned epoch = 0;
  uint32_t sparity = 0, lparity = 0;

  // phase timers (PLAS_DEBUG): 0 bookkeeping, 1..4 GEMM phases (incl. barrier), 5 query load, 6 score compute, 7 score exchange,
  // 8 softmax + logits push, 9 context, 10 argmax + cell 0, 11 attention barrier wait; fine[] = GEMM phase 0 detail by the role
  // threads (time since the phase started): 0 first tile landed, 1 MMAs issued, 2 accumulator ready, 3 accumulator read, 4 epilogue done
  unsigned long long tacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  unsigned long long tfine[5] = {0, 0, 0, 0, 0};
  unsigned long long tlast = 0;
  __shared__ unsigned long long s_tphase;
  const bool timing = p.dbg != nullptr && blockIdx.x == 0 && tid == 0;
  const bool fine = p.dbg != nullptr && blockIdx.x == 0;
  auto stamp = [&](int slot) {
    if (timing) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      tacc[slot] += now - tlast;
      tlast = now;
      *(volatile unsigned long long*)&s_tphase = now;
    }
  };
  auto fine_stamp = [&](int slot) {
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    tfine[slot] += now - *(volatile unsigned long long*)&s_tphase;
  };
  if (timing) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tlast));

  // one [128 x Ud] x [Ud x ncols] product: copy producer (warp 1: one contiguous bulk copy per k-block tile of h), MMA issuer
  // (warp 0), result in TMEM columns 0..ncols-1.  The tiles go through the ring in waves of NSTA: all copies of a wave are
  // issued back to back (normally ONE wave: a stage per k block), a later wave waits until the MMAs of the previous one have
  // read their stages.  All CTAs read the SAME tiles; walking them from a per-CTA offset keeps the CTAs from hammering the
  // same few L2 lines in lock step.
  auto gemm_phase = [&](const unsigned char* a_tiles, uint32_t w_smem, int ncols, bool detail) {
    const int rot = (int)(((long long)blockIdx.x * nkb) / gridDim.x) % nkb;
    const uint32_t idesc = umma_idesc_bf16(m64 ? 64 : 128, ncols);
    if (warp == 1) {
      if (elect_one()) {
        for (int i0 = 0; i0 < nkb; i0 += NSTA) {
          if (i0 > 0) {
            mbar_wait(wave_done, wave_parity);
            wave_parity ^= 1u;
          }
          const int n = min(NSTA, nkb - i0);
          int kb = rot + i0;
          if (kb >= nkb) kb -= nkb;
          for (int j = 0; j < n; ++j) {
            mbar_expect_tx(fullA(j), (uint32_t)STA);
            df_bulk_g2s(ring + j * STA, a_tiles + (size_t)kb * STA, (uint32_t)STA, fullA(j));
            if (++kb == nkb) kb = 0;
          }
        }
      }
      __syncwarp();
    } else if (warp == 0) {
      if (elect_one()) {
        tc_fence_after();
        // lean issue loop: descriptors are base + a multiple of the tile size (the single issuing thread is the critical path of
        // the phase: rebuilding them from addresses, or a modulo by the run-time k-block count, costs more than the MMAs)
        const uint64_t adesc0 = umma_smem_desc(ring), bdesc0 = umma_smem_desc(w_smem);
        const uint32_t astep = (uint32_t)STA >> 4, bstep = (uint32_t)(ncols * 128) >> 4;
        for (int i0 = 0; i0 < nkb; i0 += NSTA) {
          const int n = min(NSTA, nkb - i0);
          int kb = rot + i0;
          if (kb >= nkb) kb -= nkb;
          for (int j = 0; j < n; ++j) {
            mbar_wait(fullA(j), (full_bits >> j) & 1u);
            full_bits ^= 1u << j;
            if (detail && i0 + j == 0) fine_stamp(0);
            tc_fence_after();
            const uint64_t adesc = adesc0 + (uint64_t)(j * astep);
            const uint64_t bdesc = bdesc0 + (uint64_t)(kb * bstep);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (i0 | j | k) != 0 ? 1u : 0u);
            if (++kb == nkb) kb = 0;
          }
          if (i0 + n < nkb) umma_commit(wave_done);
        }
        umma_commit(tfull);
        if (detail) fine_stamp(1);
      }
      __syncwarp();
    }
  };

  const int n_items = AS * B;
  const int Wh = W4 / AS;                        // zctx columns of a part
  const int ncg = Wh / 8;                        // 8-column groups (16 bytes of VW) of a part
  const int NPR = 256 / ncg;                     // row groups that split the memory rows of the context sum
  const int cgi = tid % ncg, pri = tid / ncg;    // thread -> (column group, row group); pri >= NPR idles
  const bool ctx_thread = pri < NPR;
  const bool cell_thread = ctx_thread && pri == 0;
  const int col0 = part * Wh + cgi * 8;          // first of the 8 gate columns (two hidden units) a cell_thread owns

  // ---- cell 0 for gate columns col0..col0+7 of utterance b: z = zctx + (h0 . W0h + b0 in zb) + W_emb[id] ----
  auto cell0 = [&](int b, const float* zctx, const float* zb, float2 cprev, int id, int par_out) {
    const int idc = max(0, min(id, V - 1));
    const uint4 e = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(d.w_emb) + (size_t)idc * W4 + col0));
    const unsigned ew[4] = {e.x, e.y, e.z, e.w};
    float z[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      z[2 * j] = zctx[2 * j] + zb[2 * j] + __uint_as_float(ew[j] << 16);
      z[2 * j + 1] = zctx[2 * j + 1] + zb[2 * j + 1] + __uint_as_float(ew[j] & 0xffff0000u);
    }
    const int u0 = col0 >> 2;
    float c_a, h_a, c_b, h_b;
    lstm_gates_fast(z[0], z[1], z[2], z[3], cprev.x, c_a, h_a);
    lstm_gates_fast(z[4], z[5], z[6], z[7], cprev.y, c_b, h_b);
    *reinterpret_cast<float2*>(p.c0 + (size_t)b * Ud + u0) = make_float2(c_a, c_b);
    *reinterpret_cast<__nv_bfloat162*>(p.hbuf[0] + (size_t)par_out * hpar + df_h_off(b, u0, STA)) = __floats2bfloat162_rn(h_a, h_b);
  };
  // the operands of cell 0 that do not depend on the sampled id: h0 . W0h (ZH0) + b0, and the cell state
  auto cell0_operands = [&](int b, float* zb, float2& cprev, bool first) {
    const float4* bb = reinterpret_cast<const float4*>(d.b_cell[0] + col0);
    const float4 b_a = __ldg(bb), b_b = __ldg(bb + 1);
    float4 zh_a = make_float4(0.f, 0.f, 0.f, 0.f), zh_b = zh_a;
    cprev = make_float2(0.f, 0.f);
    if (!first) {
      const float4* zh = reinterpret_cast<const float4*>(p.zh0 + (size_t)b * W4 + col0);
      zh_a = __ldcg(zh); zh_b = __ldcg(zh + 1);
      cprev = *reinterpret_cast<const float2*>(p.c0 + (size_t)b * Ud + (col0 >> 2));
    }
    zb[0] = zh_a.x + b_a.x; zb[1] = zh_a.y + b_a.y; zb[2] = zh_a.z + b_a.z; zb[3] = zh_a.w + b_a.w;
    zb[4] = zh_b.x + b_b.x; zb[5] = zh_b.y + b_b.y; zb[6] = zh_b.z + b_b.z; zb[7] = zh_b.w + b_b.w;
  };

  // ---- prologue: cell 0 of step 0 (no attention yet: zctx = 0, h0 = 0, c0 = 0; input = sos / the first forced id) ----
  if (max_iter > 0) {
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int b = item / AS;
      if (cell_thread) {
        const float zero[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float zb[8];
        float2 cprev;
        cell0_operands(b, zb, cprev, true);
        const int id0 = d.teacher_forced ? d.forced_ids[(size_t)b * d.max_steps] : d.sos_id;
        cell0(b, zero, zb, cprev, id0, 0);
      }
    }
    df_grid_barrier(p.bar, epoch);
  }

  int t = 0;
  while (true) {
    // ---- result of step t-1: argmax, finished / sequence-length logic (replicated), outputs (CTA 0)
    if (t > 0 && row_valid) {
      const int bi = __ldcg(p.next_ids + row);  // argmax of step t-1, written by the utterance's attention CTAs
      if (blockIdx.x == 0) d.sample_ids[(size_t)row * d.max_steps + (t - 1)] = bi;
      if (!d.teacher_forced) {
        if (!finished && blockIdx.x == 0) d.seq_len[row] = t;
        finished = finished || (bi == d.eos_id) || (t >= max_iter);
      }
    }
    if (t >= max_iter) break;
    {
      const int all_done = __syncthreads_and(row_valid ? finished : 1);
      if (!d.teacher_forced && t > 0 && all_done) break;
    }
    const int par = t & 1;
    stamp(0);

    // ---------------- GEMM phases: A = h_ph (K = Ud) ----------------
#pragma unroll
    for (int ph = 0; ph < L; ++ph) {
      const int ng = n_groups(ph);
      gemm_phase(p.hbuf[ph] + (size_t)par * hpar, base + (uint32_t)p.off_w[ph], ng * 16, fine && ph == 0);
      if (row_thread) {
        // one lane per warp polls: 128 threads spinning on the mbarrier would queue ahead of the producer's and the copy
        // engine's own barrier operations in the shared-memory atomic unit
        if (lane == 0) mbar_wait(tfull, acc_parity);
        __syncwarp();
        if (fine && ph == 0 && tid == 128) fine_stamp(2);
        tc_fence_after();
        const uint32_t trow = tmem_base + ((uint32_t)((warp - 4) * 32) << 16);
        uint32_t acc[3][16];
        tmem_ld16(trow, acc[0]);
        if (ng > 1) tmem_ld16(trow + 16u, acc[1]);
        if (ng > 2) tmem_ld16(trow + 32u, acc[2]);
        tmem_ld_wait();
        if (fine && ph == 0 && tid == 128) fine_stamp(3);
        if (row_valid) {
          int g = 0;
          if (has_x(ph)) {  // layer ph+1: z = h_ph . W_x + (h_{ph+1,t-1} . W_h carried from the previous step) + b
            const int l1 = ph + 1 < L ? ph + 1 : L - 1;
            __nv_bfloat16 hq[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float cn, hn;
              lstm_gates_fast(__uint_as_float(acc[0][4 * u]) + carry[l1][4 * u] + s_bias[l1 * 16 + 4 * u],
                              __uint_as_float(acc[0][4 * u + 1]) + carry[l1][4 * u + 1] + s_bias[l1 * 16 + 4 * u + 1],
                              __uint_as_float(acc[0][4 * u + 2]) + carry[l1][4 * u + 2] + s_bias[l1 * 16 + 4 * u + 2],
                              __uint_as_float(acc[0][4 * u + 3]) + carry[l1][4 * u + 3] + s_bias[l1 * 16 + 4 * u + 3],
                              c_state[l1][u], cn, hn);
              c_state[l1][u] = cn;
              hq[u] = __float2bfloat16_rn(hn);
            }
            *reinterpret_cast<uint2*>(p.hbuf[l1] + (size_t)par * hpar + df_h_off(row, slice * 4, STA)) = *reinterpret_cast<const uint2*>(hq);
            g = 1;
          }
          if (ph == 0) {  // h0_t . W0h: read by the attention CTAs when they run cell 0 of step t+1
            float4* zd = reinterpret_cast<float4*>(p.zh0 + (size_t)row * W4 + slice * 16);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              zd[j] = make_float4(__uint_as_float(acc[g][4 * j]), __uint_as_float(acc[g][4 * j + 1]), __uint_as_float(acc[g][4 * j + 2]),
                                  __uint_as_float(acc[g][4 * j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) carry[ph][j] = __uint_as_float(acc[g][j]);
          }
          if (has_q(ph)) {
            float4* qd = reinterpret_cast<float4*>(p.qbuf + (size_t)row * Ud + slice * 16);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              qd[j] = make_float4(__uint_as_float(acc[g + 1][4 * j]), __uint_as_float(acc[g + 1][4 * j + 1]),
                                  __uint_as_float(acc[g + 1][4 * j + 2]), __uint_as_float(acc[g + 1][4 * j + 3]));
          }
        }
        tc_fence_before();
        if (fine && ph == 0 && tid == 128) fine_stamp(4);
      }
      acc_parity ^= 1u;
      // the barrier after the top phase is only needed when its result is read in this step (query layer) or by the
      // cell 0 that follows the attention (L == 1: ZH0)
      if (ph < L - 1 || qlayer || L == 1) {
        df_grid_barrier(p.bar, epoch);
      } else {
        // the attention phase reuses the ring region, also for what the partner CTAs push into it: every CTA of the
        // cluster must be done with its MMAs first
        tc_fence_before();
        asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
      }
      stamp(1 + ph);
    }

    // ---------------- attention + logits + cell 0 of step t+1 ----------------
    {
      const unsigned char* Htop = p.hbuf[L - 1] + (size_t)par * hpar;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int b = item / AS;
        const int len = min(d.mem_len[b], Tm);
        // ---- query ----
        for (int u = tid; u < Ud; u += 256)
          s_q[u] = bahdanau ? __ldcg(p.qbuf + (size_t)b * Ud + u)
                   : custom ? fmaxf(__ldcg(p.qbuf + (size_t)b * Ud + u), 0.f)
                            : __uint_as_float((unsigned)__ldcg(reinterpret_cast<const unsigned short*>(Htop + df_h_off(b, u, STA))) << 16);
        // the id-independent operands of cell 0 (t+1): issued now, consumed after the argmax
        float zb[8];
        float2 cprev = make_float2(0.f, 0.f);
        if (cell_thread) cell0_operands(b, zb, cprev, false);
        // PV rows of this part (f32 [rows][V]) are prefetched into the idle ring region with cp.async while scores run
        const int rchunk = (Tm + AS - 1) / AS;
        const int r_lo = min(part * rchunk, len), r_hi = min(r_lo + rchunk, len);
        const bool pv_smem = (V % 4 == 0) && (size_t)rchunk * V * 4 <= (size_t)p.pv_cap;
        if (pv_smem) {
          const float* pvb = d.pv + ((size_t)b * Tm + r_lo) * d.pv_ld;
          const int v4 = V / 4;
          const uint32_t pv_u = base + (uint32_t)p.off_pv;
          for (int i = tid; i < (r_hi - r_lo) * v4; i += 256) {
            const int rr = i / v4, c = i - rr * v4;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(pv_u + (uint32_t)((rr * V + 4 * c) * 4)),
                         "l"(pvb + (size_t)rr * d.pv_ld + 4 * c) : "memory");
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
        }
        df_sync();
        stamp(5);
        // ---- scores over this part's key depth: warp per memory row, lane owns 16-byte chunks lane (, lane+32) ----
        if (len > 0) {
          if (n_c8 <= 64) {
            float qreg[16], vreg[16];
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
              const int c8 = lane + 32 * cc;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                qreg[cc * 8 + j] = (c8 < n_c8) ? s_q[part * Dk + c8 * 8 + j] : 0.f;
                vreg[cc * 8 + j] = (c8 < n_c8 && bahdanau) ? s_v[part * Dk + c8 * 8 + j] : 0.f;
              }
            }
            const uint4* kres = reinterpret_cast<const uint4*>(smem + p.off_keys);
            const uint4* kglb = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(d.keys) + (size_t)b * Tm * Ud + part * Dk);
            const bool two = n_c8 > 32;
            if (p.keys_res) {
              if (bahdanau) { if (two) df_scores<2, true, true>(kres, n_c8, n_c8, len, warp, lane, qreg, vreg, s_score);
                              else df_scores<1, true, true>(kres, n_c8, n_c8, len, warp, lane, qreg, vreg, s_score); }
              else { if (two) df_scores<2, false, true>(kres, n_c8, n_c8, len, warp, lane, qreg, vreg, s_score);
                     else df_scores<1, false, true>(kres, n_c8, n_c8, len, warp, lane, qreg, vreg, s_score); }
            } else {
              if (bahdanau) { if (two) df_scores<2, true, false>(kglb, Ud / 8, n_c8, len, warp, lane, qreg, vreg, s_score);
                              else df_scores<1, true, false>(kglb, Ud / 8, n_c8, len, warp, lane, qreg, vreg, s_score); }
              else { if (two) df_scores<2, false, false>(kglb, Ud / 8, n_c8, len, warp, lane, qreg, vreg, s_score);
                     else df_scores<1, false, false>(kglb, Ud / 8, n_c8, len, warp, lane, qreg, vreg, s_score); }
            }
          } else {  // very wide decoders: chunks straight from memory
            const __nv_bfloat16* keys = reinterpret_cast<const __nv_bfloat16*>(d.keys) + (size_t)b * Tm * Ud + part * Dk;
            for (int r = warp; r < len; r += 8) {
              float acc = 0.f;
              const uint4* kr = reinterpret_cast<const uint4*>(keys + (size_t)r * Ud);
              for (int c8 = lane; c8 < n_c8; c8 += 32) {
                const uint4 k4 = __ldg(kr + c8);
                const unsigned kw[4] = {k4.x, k4.y, k4.z, k4.w};
                const int u0 = part * Dk + c8 * 8;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float k0 = __uint_as_float(kw[j] << 16), k1 = __uint_as_float(kw[j] & 0xffff0000u);
                  if (bahdanau) {
                    acc = fmaf(s_v[u0 + 2 * j], df_tanh_mufu(k0 + s_q[u0 + 2 * j]), acc);
                    acc = fmaf(s_v[u0 + 2 * j + 1], df_tanh_mufu(k1 + s_q[u0 + 2 * j + 1]), acc);
                  } else {
                    acc = fmaf(k0, s_q[u0 + 2 * j], acc);
                    acc = fmaf(k1, s_q[u0 + 2 * j + 1], acc);
                  }
                }
              }
              acc = warp_sum(acc);
              if (lane == 0) s_score[r] = acc;
            }
          }
        }
        df_sync();
        stamp(6);
        // ---- swap partial scores with the other parts of the utterance (DSMEM) and add them in part order ----
        {
          if (tid == 0) mbar_expect_tx(sbar, (uint32_t)((AS - 1) * p.tm_pad * 4));
          for (int q = 0; q < AS; ++q) {
            if (q == part) continue;
            const uint32_t dst0 = df_mapa(smem_u32(s_peer + (size_t)part * p.tm_pad), (uint32_t)(pbase + q));
            const uint32_t rb = df_mapa(sbar, (uint32_t)(pbase + q));
            for (int i = tid; i < p.tm_pad / 4; i += 256) {
              const float4 v = *reinterpret_cast<const float4*>(s_score + 4 * i);
              df_st_async_v4(dst0 + 16 * i, make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w)), rb);
            }
          }
          if (tid == 0) mbar_wait(sbar, sparity);
          sparity ^= 1u;
          df_sync();  // the partners' scores are in, and every thread has read its chunks of the own ones before they are overwritten
          for (int tm = tid; tm < len; tm += 256) {
            float s = 0.f;
            for (int q = 0; q < AS; ++q) s += (q == part) ? s_score[tm] : s_peer[(size_t)q * p.tm_pad + tm];
            s_score[tm] = s;
          }
          df_sync();
        }
        stamp(7);
        if (monotonic) {
          // tf.contrib.seq2seq.monotonic_attention(mode='parallel'): two prefix sums along memory time
          float* sa = s_scan;
          float* sb = s_scan + p.tm_pad;
          const float tiny = 1.17549435e-38f;
          const float* prev = p.align_state + ((size_t)(par ^ 1) * B + b) * Tm;
          for (int tm = tid; tm < Tm; tm += 256) {
            const float pc = (tm < len) ? sigmoidf_acc(s_score[tm] + d.score_bias) : 0.f;
            s_score[tm] = pc;
            sa[tm] = logf(fminf(fmaxf(1.f - pc, tiny), 1.f));
          }
          df_sync();
          df_block_scan(sa, Tm, s_red + 32);  // inclusive scan of the logs
          for (int tm = tid; tm < Tm; tm += 256) {
            const float cpv = expf(tm > 0 ? sa[tm - 1] : 0.f);  // exclusive cumulative product of (1 - p)
            const float pv = (t == 0) ? (tm == 0 ? 1.f : 0.f) : __ldcg(prev + tm);
            sb[tm] = pv / fminf(fmaxf(cpv, 1e-10f), 1.f);
            s_score[tm] = s_score[tm] * cpv;  // p * cp
          }
          df_sync();
          df_block_scan(sb, Tm, s_red + 32);
          float* ra = sb;
          for (int tm = tid; tm < Tm; tm += 256) s_score[tm] = s_score[tm] * ra[tm];
          df_sync();
        } else {
          float m = -INFINITY;
          for (int tm = tid; tm < len; tm += 256) m = fmaxf(m, s_score[tm]);
          m = warp_max(m);
          if (lane == 0) s_red[warp] = m;
          df_sync();
          m = s_red[0];
#pragma unroll
          for (int w = 1; w < 8; ++w) m = fmaxf(m, s_red[w]);
          float sum = 0.f;
          for (int tm = tid; tm < Tm; tm += 256) {
            const float e = (tm < len) ? expf(s_score[tm] - m) : 0.f;
            s_score[tm] = e;
            sum += e;
          }
          sum = warp_sum(sum);
          if (lane == 0) s_red[8 + warp] = sum;
          df_sync();
          sum = 0.f;
#pragma unroll
          for (int w = 0; w < 8; ++w) sum += s_red[8 + w];
          for (int tm = tid; tm < Tm; tm += 256) s_score[tm] = s_score[tm] / sum;
          df_sync();
        }
        if (part == 0) {
          if (monotonic)  // read by every part of the utterance in the next step
            for (int tm = tid; tm < Tm; tm += 256) __stcg(p.align_state + ((size_t)par * B + b) * Tm + tm, s_score[tm]);
          if (d.alignment) {
            float* ar = d.alignment + ((size_t)b * d.max_steps + t) * Tm;
            for (int tm = tid; tm < Tm; tm += 256) ar[tm] = s_score[tm];
          }
        }
        // ---- partial logits over this part's memory rows: a[r_lo..r_hi) . PV[b][r_lo..r_hi); pushed to the partners ----
        {
          const int vi = tid & 63, g = tid >> 6;
          const float* pvb = d.pv + (size_t)b * Tm * d.pv_ld;
          if (pv_smem) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            df_sync();
          }
          for (int vb = 0; vb < V; vb += 64) {
            const int v = vb + vi;
            float acc = 0.f;
            if (v < V) {
              if (pv_smem) {
#pragma unroll 8
                for (int tm = r_lo + g; tm < r_hi; tm += 4) acc = fmaf(s_score[tm], s_pv[(tm - r_lo) * V + v], acc);
              } else {
                const float* pc = pvb + v;
#pragma unroll 8
                for (int tm = r_lo + g; tm < r_hi; tm += 4) acc = fmaf(s_score[tm], __ldg(pc + (size_t)tm * d.pv_ld), acc);
              }
            }
            s_lp[g * 64 + vi] = acc;
            df_sync();
            if (tid < 64 && vb + tid < p.v_pad)
              s_lmine[vb + tid] = (vb + tid < V) ? ((s_lp[tid] + s_lp[64 + tid]) + (s_lp[128 + tid] + s_lp[192 + tid])) : 0.f;
            df_sync();
          }
          if (tid == 0) mbar_expect_tx(lbar, (uint32_t)((AS - 1) * p.v_pad * 4));
          for (int i = tid; i < p.v_pad / 4; i += 256) {
            const float4 v = *reinterpret_cast<const float4*>(s_lmine + 4 * i);
            *reinterpret_cast<float4*>(s_lpart + (size_t)part * p.v_pad + 4 * i) = v;
            for (int q = 0; q < AS; ++q) {
              if (q == part) continue;
              const uint32_t dst = df_mapa(smem_u32(s_lpart + (size_t)part * p.v_pad + 4 * i), (uint32_t)(pbase + q));
              df_st_async_v4(dst, make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w)),
                             df_mapa(lbar, (uint32_t)(pbase + q)));
            }
          }
        }
        stamp(8);
        // ---- zctx over this part's columns: thread = (8-column group, row group); VW streams from L2 ----
        float zacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (ctx_thread && len > 0) {
          const uint4* vp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(d.vw) + (size_t)b * Tm * W4 + part * Wh) + cgi;
          const size_t vstride = (size_t)(W4 / 8);
          auto load_vals = [&](int r0, uint4* vv) {  // unconditional (row clamped): rows >= len get weight 0 below
#pragma unroll
            for (int i = 0; i < 8; ++i) vv[i] = __ldg(vp + (size_t)min(r0 + NPR * i, len - 1) * vstride);
          };
          auto accum = [&](int r0, const uint4* vv) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = r0 + NPR * i;
              const float a = (r < len) ? s_score[r] : 0.f;
              const unsigned w4[4] = {vv[i].x, vv[i].y, vv[i].z, vv[i].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                zacc[2 * e] = fmaf(a, __uint_as_float(w4[e] << 16), zacc[2 * e]);
                zacc[2 * e + 1] = fmaf(a, __uint_as_float(w4[e] & 0xffff0000u), zacc[2 * e + 1]);
              }
            }
          };
          // 8 x 16 bytes in flight per thread while the previous 8 rows are accumulated
          uint4 va[8], vb2[8];
          load_vals(pri, va);
          for (int r0 = pri; r0 < len; r0 += 16 * NPR) {
            load_vals(r0 + 8 * NPR, vb2);
            accum(r0, va);
            load_vals(r0 + 16 * NPR, va);
            accum(r0 + 8 * NPR, vb2);
          }
          if (pri > 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) s_comb[(size_t)(pri - 1) * Wh + cgi * 8 + i] = zacc[i];
          }
        }
        df_sync();
        if (cell_thread) {
          for (int q = 1; q < NPR; ++q) {
#pragma unroll
            for (int i = 0; i < 8; ++i) zacc[i] += s_comb[(size_t)(q - 1) * Wh + cgi * 8 + i];
          }
        }
        stamp(9);
        // ---- logits = sum of the parts' partial logits (part order) + bias; argmax, lowest index wins ties ----
        {
          float best = -INFINITY;
          int bi = 0x7fffffff;
          if (tid < 64) {
            if (lane == 0) mbar_wait(lbar, lparity);
            __syncwarp();
            for (int vb = 0; vb < V; vb += 64) {
              const int v = vb + tid;
              if (v < V) {
                float sv = 0.f;
                for (int q = 0; q < AS; ++q) sv += s_lpart[(size_t)q * p.v_pad + v];
                sv += d.b_proj[v];
                if (part == 0) d.logits[((size_t)b * d.max_steps + t) * V + v] = sv;
                if (sv > best) { best = sv; bi = v; }
              }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const float ob = __shfl_xor_sync(0xffffffffu, best, o);
              const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
              if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (lane == 0) { s_red[16 + 2 * warp] = best; s_red[17 + 2 * warp] = __int_as_float(bi); }
          }
          lparity ^= 1u;
          df_sync();
          float b0 = s_red[16], b1 = s_red[18];
          int i0 = __float_as_int(s_red[17]), i1 = __float_as_int(s_red[19]);
          if (b1 > b0 || (b1 == b0 && i1 < i0)) { b0 = b1; i0 = i1; }
          const int id_t = i0 == 0x7fffffff ? 0 : i0;
          if (part == 0 && tid == 0) __stcg(p.next_ids + b, id_t);
          // ---- cell 0 of step t+1 for this part's hidden units ----
          if (t + 1 < max_iter && cell_thread) {
            const int id_in = d.teacher_forced ? d.forced_ids[(size_t)b * d.max_steps + t + 1] : id_t;
            cell0(b, zacc, zb, cprev, id_in, par ^ 1);
          }
          df_sync();  // s_red / s_score / s_lpart are rewritten by the next item
        }
        stamp(10);
      }
      df_grid_barrier(p.bar, epoch);
      stamp(11);
    }
    ++t;
  }
  if (timing)
    for (int i = 0; i < 12; ++i) p.dbg[i] = tacc[i];
  if (fine) {  // written by the role threads that took the stamps
    if (warp == 0 && tfine[0]) { p.dbg[12] = tfine[0]; p.dbg[13] = tfine[1]; }
    if (tid == 128) { p.dbg[14] = tfine[2]; p.dbg[15] = tfine[3]; p.dbg[16] = tfine[4]; }
  }
  if (blockIdx.x == 0 && tid == 0) *d.n_steps = t;
  tc_fence_before();
  __syncthreads();
  // no CTA may exit while a cluster peer can still push into its shared memory or read its barriers
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem_base) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct DecFoldPlan {
  bool ok;
  int as, keys_res, n_stages_a, stage_a, tm_pad, v_pad;
  int off_w[4], off_keys, off_ring, off_comb, off_pv, pv_cap, off_att, off_misc;
  size_t smem;
  size_t ws_off_h[4], ws_off_q, ws_off_zh0, ws_off_c0, ws_off_align, ws_off_ids, ws_off_bar, ws_off_dbg, ws_total;
};

static DecFoldPlan dec_fold_plan(const plas_dec_desc& d) {
  DecFoldPlan pl;
  memset(&pl, 0, sizeof(pl));
  const bool bahdanau = d.attention_type == PLAS_ATT_BAHDANAU || d.attention_type == PLAS_ATT_CUSTOM;  // a query layer in the top phase
  const bool shape_ok = d.dtype == PLAS_BF16 && d.vw && d.pv && d.w_h_tc[0] && d.B >= 1 && d.B <= 128 && d.Ud % 64 == 0 &&
                        d.Ud >= 64 && (d.Ud / 4) <= num_sms() && d.Tm <= 4096 && d.n_layers >= 1 && d.n_layers <= 4 &&
                        ((d.attention_type >= PLAS_ATT_LUONG && d.attention_type <= PLAS_ATT_LUONG_MONOTONIC) ||
                         d.attention_type == PLAS_ATT_CUSTOM) &&
                        (!bahdanau || d.w_query_tc);
  if (!shape_ok) return pl;
  for (int l = 1; l < d.n_layers; ++l)
    if (!d.w_x_tc[l] || !d.w_h_tc[l]) return pl;
  const int L = d.n_layers, Ud = d.Ud, nkb = Ud / 64, grid = Ud / 4;
  pl.as = (4 * d.B <= grid) ? 4 : 2;
  // zctx column groups of a part must fit the 256 threads; key depth of a part in whole 16-byte chunks
  if ((4 * Ud / pl.as) % 8 != 0 || (4 * Ud / pl.as) / 8 > 256 || (Ud / pl.as) % 8 != 0) return pl;
  pl.tm_pad = (d.Tm + 3) & ~3;
  pl.v_pad = (d.V + 3) & ~3;
  int off = 0;
  for (int ph = 0; ph < L; ++ph) {
    const int ng = (ph < L - 1 ? 1 : 0) + 1 + ((ph == L - 1 && bahdanau) ? 1 : 0);  // the q group exists on the first Ud/16 CTAs only; sized for them
    pl.off_w[ph] = off;
    off += ng * 32 * Ud;  // 16 columns x Ud x bf16 per group
  }
  pl.stage_a = d.B <= 64 ? 8192 : 16384;  // one k-block tile of h: 64 or 128 rows x 128 B
  const int misc = 256 + 4 * (64 + Ud) + 64;  // barriers, TMEM slot, biases, attention_v
  int avail = 227 * 1024 - 1024 - off - misc;
  // the region the phases share: activation ring of the GEMM phases (ideally a stage per k block: one wave of copies) |
  // attention phase: zctx partials (8 KB), the PV window of a part when it fits, scratch arrays
  const int att_bytes = (4 * (64 + 256 + 5 * pl.v_pad + Ud + 7 * pl.tm_pad) + 1023) & ~1023;
  const int rchunk = (d.Tm + pl.as - 1) / pl.as;
  const int pv_bytes = (d.V % 4 == 0) ? ((rchunk * d.V * 4 + 1023) & ~1023) : 0;
  int att_need = 8192 + att_bytes;
  if (att_need > avail) return pl;
  int pv_win = 0;
  if (pv_bytes && att_need + pv_bytes <= avail) { pv_win = pv_bytes; att_need += pv_bytes; }
  int ns = nkb < DF_MAX_STAGES ? nkb : DF_MAX_STAGES;
  // the resident key slice (one attention item per CTA) comes before the last ring stages: give up stages (never below
  // two, i.e. more waves) only if that is what lets the keys fit
  pl.keys_res = 0;
  int key_bytes = 0;
  {
    const int kb = ((d.Tm * (Ud / pl.as) * 2) + 1023) & ~1023;
    const bool single_item = pl.as * d.B <= grid;
    const char* e = getenv("PLAS_DEC_KEYS_RES");
    if (single_item && !(e && atoi(e) == 0)) {
      int ns_k = ns;
      while (ns_k > 2 && (ns_k * pl.stage_a > att_need ? ns_k * pl.stage_a : att_need) + kb > avail) --ns_k;
      if ((ns_k * pl.stage_a > att_need ? ns_k * pl.stage_a : att_need) + kb <= avail && ns_k * 2 >= ns) {
        pl.keys_res = 1;
        key_bytes = kb;
        ns = ns_k;
      }
    }
  }
  while (ns > 1 && (ns * pl.stage_a > att_need ? ns * pl.stage_a : att_need) + key_bytes > avail) --ns;
  const int region = ns * pl.stage_a > att_need ? ns * pl.stage_a : att_need;
  if (region + key_bytes > avail) return pl;
  pl.n_stages_a = ns;
  pl.off_ring = off;
  pl.off_comb = off;
  pl.off_pv = off + 8192;
  pl.pv_cap = pv_win;
  pl.off_att = off + 8192 + pv_win;
  off += region;
  if (pl.keys_res) {
    pl.off_keys = off;
    off += key_bytes;
  }
  pl.off_misc = off;
  pl.smem = (size_t)off + misc + 1024;
  // 64-row tiles: the M=128 MMA reads 8 KB past a stage (rows 64..127 feed TMEM lanes nobody reads); keep those bytes allocated
  const size_t over = (size_t)pl.off_ring + (size_t)pl.n_stages_a * pl.stage_a + 8192 + 1024;
  if (d.B <= 64 && pl.smem < over) pl.smem = over;
  if (pl.smem > 227 * 1024) return pl;
  size_t w = 0;
  auto take = [&](size_t bytes) { size_t o = w; w += (bytes + 255) & ~size_t(255); return o; };
  for (int l = 0; l < 4; ++l) pl.ws_off_h[l] = take(l < L ? 2 * (size_t)nkb * pl.stage_a : 0);
  pl.ws_off_q = take((size_t)d.B * Ud * 4);
  pl.ws_off_zh0 = take((size_t)d.B * 4 * Ud * 4);
  pl.ws_off_c0 = take((size_t)d.B * Ud * 4);
  pl.ws_off_align = take(2 * (size_t)d.B * d.Tm * 4);
  pl.ws_off_ids = take((size_t)d.B * 4);
  pl.ws_off_bar = take(4);
  pl.ws_off_dbg = take(512);
  pl.ws_total = w;
  pl.ok = true;
  return pl;
}

size_t dec_fold_workspace_bytes(const plas_dec_desc& d) {
  const DecFoldPlan pl = dec_fold_plan(d);
  return pl.ok ? pl.ws_total : 0;
}

template <int L>
static cudaError_t dec_fold_launch_l(const DecFoldArgs& a, int grid, size_t smem, cudaStream_t stream, int* max_clusters) {
  cudaError_t e = cudaFuncSetAttribute(decoder_fold_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(DF_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = DF_CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // the grid barrier needs every cluster co-resident
  e = cudaOccupancyMaxActiveClusters(max_clusters, decoder_fold_kernel<L>, &cfg);
  if (e != cudaSuccess) return e;
  if (*max_clusters * DF_CL < grid) return cudaErrorCooperativeLaunchTooLarge;
  return cudaLaunchKernelEx(&cfg, decoder_fold_kernel<L>, a);
}

// Returns PLAS_OK after a launch, 1 when the shape is not eligible (caller falls back), <0 on error.
int dec_fold_launch(const plas_dec_desc& d, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  const char* force = getenv("PLAS_DEC_IMPL");
  if (force && (strcmp(force, "simt") == 0 || strcmp(force, "tc1") == 0)) return 1;
  const DecFoldPlan pl = dec_fold_plan(d);
  if (!pl.ok) return 1;
  PLAS_REQUIRE(workspace_bytes >= pl.ws_total, "decoder(fold): workspace %zu < %zu", workspace_bytes, pl.ws_total);
  PLAS_CUDA(cudaMemsetAsync(workspace, 0, pl.ws_total, stream));
  PLAS_CUDA(cudaMemsetAsync(d.seq_len, 0, (size_t)d.B * 4, stream));
  PLAS_CUDA(cudaMemsetAsync(d.n_steps, 0, 4, stream));
  if (d.max_steps == 0) return PLAS_OK;

  DecFoldArgs a;
  memset(&a, 0, sizeof(a));
  a.d = d;
  unsigned char* ws = (unsigned char*)workspace;
  for (int l = 0; l < 4; ++l) {
    a.hbuf[l] = ws + pl.ws_off_h[l];
    a.w_x[l] = (const unsigned char*)d.w_x_tc[l];
    a.w_h[l] = (const unsigned char*)d.w_h_tc[l];
    a.off_w[l] = pl.off_w[l];
  }
  a.w_q = (const unsigned char*)d.w_query_tc;
  a.qbuf = (float*)(ws + pl.ws_off_q);
  a.zh0 = (float*)(ws + pl.ws_off_zh0);
  a.c0 = (float*)(ws + pl.ws_off_c0);
  a.align_state = (float*)(ws + pl.ws_off_align);
  a.next_ids = (int*)(ws + pl.ws_off_ids);
  a.bar = (unsigned*)(ws + pl.ws_off_bar);
  a.dbg = getenv("PLAS_DEBUG") ? (unsigned long long*)(ws + pl.ws_off_dbg) : nullptr;
  a.as = pl.as;
  {
    const char* e = getenv("PLAS_DEC_M64");
    a.m64 = (d.B <= 64 && !(e && atoi(e) == 0)) ? 1 : 0;
  }
  a.keys_res = pl.keys_res;
  a.n_stages_a = pl.n_stages_a;
  a.stage_a = pl.stage_a;
  a.off_keys = pl.off_keys;
  a.off_ring = pl.off_ring;
  a.off_comb = pl.off_comb;
  a.off_pv = pl.off_pv;
  a.pv_cap = pl.pv_cap;
  a.off_att = pl.off_att;
  a.off_misc = pl.off_misc;
  a.tm_pad = pl.tm_pad;
  a.v_pad = pl.v_pad;
  const int grid = d.Ud / 4;
  int max_clusters = 0;
  cudaError_t le;
  switch (d.n_layers) {
    case 1: le = dec_fold_launch_l<1>(a, grid, pl.smem, stream, &max_clusters); break;
    case 2: le = dec_fold_launch_l<2>(a, grid, pl.smem, stream, &max_clusters); break;
    case 3: le = dec_fold_launch_l<3>(a, grid, pl.smem, stream, &max_clusters); break;
    default: le = dec_fold_launch_l<4>(a, grid, pl.smem, stream, &max_clusters); break;
  }
  if (getenv("PLAS_DEBUG"))
    fprintf(stderr, "[plas] decoder fold path: L=%d as=%d keys_res=%d grid=%d smem=%zu ring %d x %d B pv_cap=%d max_active_clusters=%d launch: %s\n",
            d.n_layers, pl.as, pl.keys_res, grid, pl.smem, pl.n_stages_a, pl.stage_a, pl.pv_cap, max_clusters, cudaGetErrorString(le));
  if (le == cudaErrorCooperativeLaunchTooLarge) {
    (void)cudaGetLastError();
    return 1;  // not all clusters fit at once: the caller falls back
  }
  PLAS_CUDA(le);
  if (a.dbg) {  // debug only: synchronises
    unsigned long long h[17];
    PLAS_CUDA(cudaStreamSynchronize(stream));
    PLAS_CUDA(cudaMemcpy(h, a.dbg, sizeof(h), cudaMemcpyDeviceToHost));
    fprintf(stderr, "[plas] decoder fold phase time (us, CTA 0): bookkeeping %.1f  gemm %.1f %.1f %.1f %.1f  query load %.1f  scores %.1f  exchange %.1f  softmax+logits %.1f  context %.1f  argmax+cell0 %.1f  barrier %.1f\n",
            h[0] / 1e3, h[1] / 1e3, h[2] / 1e3, h[3] / 1e3, h[4] / 1e3, h[5] / 1e3, h[6] / 1e3, h[7] / 1e3, h[8] / 1e3, h[9] / 1e3, h[10] / 1e3, h[11] / 1e3);
    fprintf(stderr, "[plas]   gemm phase 0 detail (us after phase start, summed): first tile %.1f  MMAs issued %.1f  accumulator ready %.1f  accumulator read %.1f  epilogue done %.1f\n",
            h[12] / 1e3, h[13] / 1e3, h[14] / 1e3, h[15] / 1e3, h[16] / 1e3);
  }
  return PLAS_OK;
}

}  // namespace plas
