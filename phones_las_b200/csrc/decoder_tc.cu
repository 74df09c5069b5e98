// K4 (tensor-core path, bf16) -- weight-stationary attention decoder for sm_100a.
//
// Same contract as decoder.cu (AttentionWrapper(MultiRNNCell) + BasicDecoder + Greedy/Training
// helper + dynamic_decode, reference las/model.py:145-349; projection utils/training_helper.py:122-153)
// but organised around what bounds a decode step on B200 -- bytes into each SM:
//
//  * LSTM layers (phase A): CTA s owns 16 gate columns (4 hidden units x i,j,f,o) of EVERY layer and
//    keeps those weight slices resident in shared memory for the whole decode, laid out as
//    K-major SWIZZLE_128B UMMA B tiles.  Per step only the activations X_l [B x K_l] stream in:
//    TMA (cp.async.bulk.tensor, 128 x 64 boxes, rows >= B zero-filled) -> 16 KB ring stages ->
//    tcgen05.mma (M=128, N=16, K=16; accumulator = 16 TMEM columns).  Epilogue thread r owns batch
//    row r: it reads its 16 pre-activations with tcgen05.ld, adds bias and the one-hot embedding
//    row, applies the TF gate math and keeps the cell state of its 4 units in registers.
//  * query layer (bahdanau): the same machinery with 16 output columns per CTA on CTAs 0..Ud/16-1.
//  * attention (phase B): work item = (utterance, half of the context channels).  A producer thread
//    streams keys[b] and values[b][:, half] through the same ring with cp.async.bulk; 8 consumer
//    warps compute scores -> masked softmax / monotonic scan -> context.  logits come from
//    PV = values x W_proj (one tcgen05 GEMM per utterance batch, fp32): logits = a . PV + b.
//  * greedy argmax / finished / sequence-length logic is replicated per CTA (deterministic), so
//    no extra grid barrier is needed for sampling; CTA 0 writes the outputs.
// Phases are separated by a grid barrier (L2 atomics); a step is L + 1 (+1 for bahdanau) barriers.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tcgen05.cuh"
#include "../../include/plas.h"

namespace plas {

constexpr int DT_THREADS = 288;   // warps 0..7 compute, warp 8 = TMA / bulk-copy producer
constexpr int DT_STAGE = 16384;   // ring stage: one 128 x 64 bf16 TMA box
constexpr int DT_MAX_STAGES = 16;

struct alignas(64) DecTcArgs {
  CUtensorMap tmX[4][2];  // per layer, per parity: X_l [B][K_l] bf16
  CUtensorMap tmQ[2];     // per parity: top-layer h [B][Ud] (strided view into xbuf[L-1])
  plas_dec_desc d;
  const unsigned char* w_tc[4];  // [Ud/4][K_l/64][2048 B] swizzled UMMA B tiles
  const unsigned char* wq_tc;    // [Ud/16][Ud/64][2048 B]
  unsigned char* xbuf[4];        // [2][B][K_l] bf16
  float* qbuf;                   // [B][Ud]
  float* align_state;            // [B][Tm]
  int* next_ids;                 // [B] argmax of the step just decoded
  unsigned* bar;
  unsigned long long* dbg;       // optional [8] phase timers (ns, summed over steps) written by CTA 0
  int n_stages;                  // 16 KB stages of the attention ring
  int n_stages_a;                // stages of the LSTM-phase ring (8 KB when B <= 64, else 16 KB)
  int stage_a;                   // bytes per LSTM-phase stage (= TMA box bytes)
  int ksplit;                    // 1, or 4: clusters of 4 CTAs split K of the LSTM products (DSMEM reduction)
  int no_pair_split;             // debugging: every attention CTA scores the full depth
  int off_part;                  // ksplit: [4 senders][128 rows][16] f32 partial sums (inside the ring's tail)
  int off_w[4], off_wq, off_ring, off_misc;  // byte offsets from the 1024-aligned smem base
  int tm_pad;
};

__device__ __forceinline__ uint32_t mapa_cl(uint32_t local_saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async_v4_cl(uint32_t raddr, const uint4 v, uint32_t rbar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(rbar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ void dt_grid_barrier(unsigned* bar, unsigned& epoch) {
  fence_proxy_async();  // generic-proxy global/shared writes of this phase vs TMA / bulk copies of the next
  __syncthreads();      // every thread's writes happen-before thread 0's release (cumulative at gpu scope)
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
  fence_proxy_async();
}

// bahdanau scores need B*Tm*Ud tanh per decode step (6.2 M at c2): one MUFU op each.  tanh.approx.f32
// has a relative error of 2^-11, an eighth of the bf16 quantisation of the keys it is applied to.
__device__ __forceinline__ float tanh_mufu(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct RingState {
  int stage;
  uint32_t phase;
  __device__ __forceinline__ void advance(int n) {
    if (++stage == n) { stage = 0; phase ^= 1u; }
  }
};

__global__ void __launch_bounds__(DT_THREADS, 1) decoder_tc_kernel(const __grid_constant__ DecTcArgs p) {
  extern __shared__ unsigned char dt_smem_raw[];
  const plas_dec_desc& d = p.d;
  const int B = d.B, Tm = d.Tm, D = d.D, Ud = d.Ud, V = d.V, L = d.n_layers;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NST = p.n_stages, NSTA = p.n_stages_a, STA = p.stage_a;
  const uint32_t raw = smem_u32(dt_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* smem = dt_smem_raw + (base - raw);
  const uint32_t ring = base + (uint32_t)p.off_ring;
  unsigned char* ring_ptr = smem + p.off_ring;

  // misc region: barriers, TMEM slot, small arrays
  unsigned char* misc = smem + p.off_misc;
  const uint32_t misc_u = base + (uint32_t)p.off_misc;
  auto fullA = [&](int s) { return misc_u + 8u * s; };
  auto emptyA = [&](int s) { return misc_u + 8u * (DT_MAX_STAGES + s); };
  const uint32_t tfull = misc_u + 8u * (4 * DT_MAX_STAGES);
  const uint32_t pbar = misc_u + 8u * (4 * DT_MAX_STAGES) + 16u;  // K-split partial sums landed (tx bytes)
  const uint32_t sbar = misc_u + 8u * (4 * DT_MAX_STAGES) + 24u;  // the pair partner's partial scores landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 8 * (4 * DT_MAX_STAGES + 1));
  float* s_bias = reinterpret_cast<float*>(misc + 1536);         // [4][16]
  float* s_red = s_bias + 64;                                    // [64]
  float* s_lp = s_red + 64;                                      // [8][32]
  float* s_q = s_lp + 256;                                       // [Ud]
  float* s_v = s_q + Ud;                                         // [Ud]
  float* s_score = s_v + Ud;                                     // [tm_pad]
  float* s_scan = s_score + p.tm_pad;                            // [2][tm_pad]
  float* s_comb = s_scan + 2 * p.tm_pad;                         // [D/2]
  float* s_peer = s_comb + D / 2;                                // [2][tm_pad] partner's partial scores

  const int nsl = Ud / 4;
  const int nq = Ud / 16;
  const bool bahdanau = d.attention_type == PLAS_ATT_BAHDANAU;
  const bool monotonic = d.attention_type == PLAS_ATT_LUONG_MONOTONIC;
  const int slice = blockIdx.x;
  const bool cell_cta = slice < nsl;
  const bool q_cta = bahdanau && slice < nq;
  int Kl[4];
  for (int l = 0; l < 4; ++l) Kl[l] = (l == 0) ? (D + Ud) : 2 * Ud;

  // ---- one-time setup: resident weight slices, barriers, TMEM ------------------------------
  const int KSP = p.ksplit;
  const int crank = blockIdx.x % KSP;         // rank inside the K-split cluster
  const int cbase = blockIdx.x - crank;       // first slice of the cluster
  if (cell_cta) {
    for (int l = 0; l < L; ++l) {
      const int nkb = Kl[l] / 64;
      uint4* dst = reinterpret_cast<uint4*>(smem + p.off_w[l]);
      if (KSP == 1) {
        const uint4* src = reinterpret_cast<const uint4*>(p.w_tc[l] + (size_t)slice * nkb * 2048);
        for (int i = tid; i < nkb * 128; i += DT_THREADS) dst[i] = __ldg(src + i);
      } else {
        // K quarter `crank` of the 4 slices of this cluster: per local k block a 64-row UMMA B tile, which in the
        // 128B-swizzled K-major layout is just the four 16-row slice tiles back to back
        const int nloc = nkb / KSP;
        for (int i = tid; i < nloc * KSP * 128; i += DT_THREADS) {
          const int w16 = i & 127, j = (i >> 7) % KSP, lkb = i / (128 * KSP);
          const uint4* src = reinterpret_cast<const uint4*>(p.w_tc[l] + ((size_t)(cbase + j) * nkb + crank * nloc + lkb) * 2048);
          dst[i] = __ldg(src + w16);
        }
      }
      if (tid < 16) s_bias[l * 16 + tid] = d.b_cell[l][slice * 16 + tid];
    }
  }
  const bool q_ksplit = KSP > 1 && (Ud / 64) % KSP == 0;
  if (q_cta) {
    const int nkb = Ud / 64;
    uint4* dst = reinterpret_cast<uint4*>(smem + p.off_wq);
    if (!q_ksplit) {
      const uint4* src = reinterpret_cast<const uint4*>(p.wq_tc + (size_t)slice * nkb * 2048);
      for (int i = tid; i < nkb * 128; i += DT_THREADS) dst[i] = __ldg(src + i);
    } else {  // same K-quarter x 4-slice layout as the cell weights
      const int nloc = nkb / KSP;
      for (int i = tid; i < nloc * KSP * 128; i += DT_THREADS) {
        const int w16 = i & 127, j = (i >> 7) % KSP, lkb = i / (128 * KSP);
        const uint4* src = reinterpret_cast<const uint4*>(p.wq_tc + ((size_t)(cbase + j) * nkb + crank * nloc + lkb) * 2048);
        dst[i] = __ldg(src + w16);
      }
    }
  }
  if (bahdanau)
    for (int u = tid; u < Ud; u += DT_THREADS) s_v[u] = d.v_att[u];
  if (tid == 0) {
    for (int s = 0; s < DT_MAX_STAGES; ++s) {
      mbar_init(fullA(s), 1);
      mbar_init(emptyA(s), 1);
    }
    mbar_init(tfull, 1);
    mbar_init(pbar, 1);
    mbar_init(sbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int l = 0; l < L; ++l)
      for (int q2 = 0; q2 < 2; ++q2) asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmX[l][q2]) : "memory");
    for (int q2 = 0; q2 < 2; ++q2) asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmQ[q2]) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();  // the resident weights were written with generic stores, UMMA reads them via the async proxy
  tc_fence_before();
  __syncthreads();
  if (KSP > 1)  // peers' mbarriers must be initialised before anyone signals them
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- replicated decode state ----------------------------------------------------------------
  const int row = tid - 128;                      // epilogue threads (warps 4..7) own batch row `row`
  const bool row_thread = tid >= 128 && tid < 256;
  const bool row_valid = row_thread && row < B;
  float c_state[4][4];
#pragma unroll
  for (int l = 0; l < 4; ++l)
#pragma unroll
    for (int u = 0; u < 4; ++u) c_state[l][u] = 0.f;
  int cur_id = d.sos_id;
  int finished = 0;

  int max_iter = d.max_steps;
  if (!d.teacher_forced) {
    int ml = 0;
    for (int b = 0; b < B; ++b) ml = max(ml, d.mem_len[b]);
    max_iter = min(max_iter, (int)rintf((float)ml * d.decoding_length_factor));
  }

  RingState prodA = {0, 0}, consA = {0, 0};
  uint32_t acc_parity = 0;
  unsigned epoch = 0;
  constexpr uint32_t IDESC = umma_idesc_bf16(128, 16);

  // optional phase timers (PLAS_DEBUG): 0 prologue, 1..4 LSTM layers (incl. barrier), 5 query, 6 attention
  unsigned long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  __shared__ unsigned long long s_tphase;          // start of the current phase (thread 0's last stamp)
  unsigned long long tfine[4] = {0, 0, 0, 0};      // layer-0 detail: first box landed, MMAs issued, accumulator ready, epilogue done
  const bool fine = p.dbg != nullptr && blockIdx.x == 0;
  unsigned long long tfineB[4] = {0, 0, 0, 0};     // attention detail: scores, softmax, context, logits done
  auto fineB = [&](int slot) {
    if (fine && tid == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      tfineB[slot] += now - *(volatile unsigned long long*)&s_tphase;
    }
  };
  auto fine_stamp = [&](int slot) {
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    tfine[slot] += now - *(volatile unsigned long long*)&s_tphase;
  };
  unsigned long long tlast = 0;
  const bool timing = p.dbg != nullptr && blockIdx.x == 0 && tid == 0;
  auto stamp = [&](int slot) {
    if (timing) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      tacc[slot] += now - tlast;
      tlast = now;
      *(volatile unsigned long long*)&s_tphase = now;
    }
  };
  if (timing) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tlast));

  // one [128 x K] x [K x 16] product: TMA producer (warp 8), MMA issuer (warp 0), result in TMEM cols 0..15
  // one [128 x (nloc*64)] x [(nloc*64) x NC] product over the k blocks kb_lo .. kb_lo+nloc-1: TMA producer (warp 8),
  // MMA issuer (warp 0), result in TMEM columns 0..NC-1.  Many CTAs read the SAME activation blocks; walking them
  // from a per-CTA offset keeps the CTAs from hammering the same few L2 lines in lock step.
  auto gemm_phase = [&](const CUtensorMap* tm, uint32_t w_smem, int kb_lo, int nloc, int ncols, uint32_t idesc, bool detail) {
    const int rot = (int)(((long long)(blockIdx.x / KSP) * nloc * KSP) / gridDim.x) % nloc;
    if (warp == 8) {
      if (elect_one()) {
        for (int i = 0; i < nloc; ++i) {
          const int lkb = (rot + i) % nloc;
          mbar_wait(emptyA(prodA.stage), prodA.phase ^ 1u);
          mbar_expect_tx(fullA(prodA.stage), (uint32_t)STA);
          tma_load_2d(ring + prodA.stage * STA, tm, (kb_lo + lkb) * 64, 0, fullA(prodA.stage));
          prodA.advance(NSTA);
        }
      }
      __syncwarp();
    } else if (warp == 0) {
      if (elect_one()) {
        tc_fence_after();
        for (int i = 0; i < nloc; ++i) {
          const int lkb = (rot + i) % nloc;
          mbar_wait(fullA(consA.stage), consA.phase);
          if (detail && i == 0) fine_stamp(0);
          tc_fence_after();
          const uint64_t adesc = umma_smem_desc(ring + consA.stage * STA);
          const uint64_t bdesc = umma_smem_desc(w_smem + (uint32_t)(lkb * ncols * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (i | k) != 0 ? 1u : 0u);
          umma_commit(emptyA(consA.stage));
          consA.advance(NSTA);
        }
        umma_commit(tfull);
        if (detail) fine_stamp(1);
      }
      __syncwarp();
    }
  };
  constexpr uint32_t IDESC64 = umma_idesc_bf16(128, 64);
  uint32_t pparity = 0, sparity = 0;
  int items_done = 0;
  float* s_part = reinterpret_cast<float*>(smem + p.off_part);
  const uint32_t s_part_u = base + (uint32_t)p.off_part;
  const uint32_t part_bytes = (uint32_t)((KSP - 1) * min(B, 128) * 64);
  const int PR = B <= 64 ? 64 : 128;  // rows of one sender's block in s_part

  // the 16 pre-activation columns of this CTA for batch row `row` (TMEM lane of the thread), after the accumulator
  // barrier: directly, or -- K-split -- own partial + the three peers' partials (DSMEM), summed in a fixed order
  auto fetch_acc16 = [&](uint32_t* r, bool ksplit) {
    const uint32_t trow = tmem_base + ((uint32_t)((warp - 4) * 32) << 16);
    if (!ksplit) {
      tmem_ld16(trow, r);
      tmem_ld_wait();
      return;
    }
    // columns 16j..16j+15 of my K-partial belong to CTA j of the cluster: keep mine, push the rest (DSMEM)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t rj[16];
      tmem_ld16(trow + (uint32_t)(16 * j), rj);
      tmem_ld_wait();
      if (j == crank) {
#pragma unroll
        for (int q4 = 0; q4 < 16; ++q4) r[q4] = rj[q4];
      } else if (row_valid) {
        const uint32_t dst = mapa_cl(s_part_u + (uint32_t)((crank * PR + row) * 64), (uint32_t)j);
        const uint32_t rb = mapa_cl(pbar, (uint32_t)j);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
          st_async_v4_cl(dst + 16 * q4, make_uint4(rj[4 * q4], rj[4 * q4 + 1], rj[4 * q4 + 2], rj[4 * q4 + 3]), rb);
      }
    }
    mbar_wait(pbar, pparity);
    pparity ^= 1u;
    if (row_valid) {  // fixed summation order (sender 0,1,2,3) so every CTA adds the same way
      float acc[16];
#pragma unroll
      for (int q4 = 0; q4 < 16; ++q4) acc[q4] = 0.f;
#pragma unroll
      for (int sr = 0; sr < 4; ++sr) {
        if (sr == crank) {
#pragma unroll
          for (int q4 = 0; q4 < 16; ++q4) acc[q4] += __uint_as_float(r[q4]);
        } else {
          const float4* pr = reinterpret_cast<const float4*>(s_part + (size_t)(sr * PR + row) * 16);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const float4 v = pr[q4];
            acc[4 * q4] += v.x; acc[4 * q4 + 1] += v.y; acc[4 * q4 + 2] += v.z; acc[4 * q4 + 3] += v.w;
          }
        }
      }
#pragma unroll
      for (int q4 = 0; q4 < 16; ++q4) r[q4] = __float_as_uint(acc[q4]);
    }
  };

  int t = 0;
  while (true) {
    // ---- result of step t-1: argmax, finished / sequence-length logic (replicated), outputs (CTA 0)
    if (t > 0 && row_valid) {
      const int bi = __ldcg(p.next_ids + row);  // argmax of step t-1, written by the utterance's attention CTA
      if (blockIdx.x == 0) d.sample_ids[(size_t)row * d.max_steps + (t - 1)] = bi;
      if (!d.teacher_forced) {
        if (!finished && blockIdx.x == 0) d.seq_len[row] = t;
        finished = finished || (bi == d.eos_id) || (t >= max_iter);
        cur_id = bi;
      }
    }
    if (t >= max_iter) break;
    {
      const int all_done = __syncthreads_and(row_valid ? finished : 1);
      if (!d.teacher_forced && t > 0 && all_done) break;
    }
    if (d.teacher_forced && row_valid) cur_id = d.forced_ids[(size_t)row * d.max_steps + t];
    const int par = t & 1;
    stamp(0);

    // ---------------- phase A: LSTM layers ----------------
    for (int l = 0; l < L; ++l) {
      if (cell_cta) {
        const int K = Kl[l];
        if (KSP == 1) gemm_phase(&p.tmX[l][par], base + (uint32_t)p.off_w[l], 0, K / 64, 16, IDESC, fine && l == 0);
        else gemm_phase(&p.tmX[l][par], base + (uint32_t)p.off_w[l], crank * (K / 64 / KSP), K / 64 / KSP, 64, IDESC64, fine && l == 0);
        if (row_thread) {
          if (KSP > 1 && tid == 128) mbar_expect_tx(pbar, part_bytes);  // partial sums of the 3 peers for my 16 columns
          // the one-hot input is a row lookup in the cell-0 kernel: fetch it while the MMAs run
          uint4 e0 = make_uint4(0u, 0u, 0u, 0u), e1 = e0;
          if (l == 0 && row_valid) {
            const int id = max(0, min(cur_id, V - 1));
            const uint4* er = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(d.w_emb) +
                                                             (size_t)id * 4 * Ud + slice * 16);
            e0 = __ldg(er);
            e1 = __ldg(er + 1);
          }
          mbar_wait(tfull, acc_parity);
          if (fine && l == 0 && tid == 128) fine_stamp(2);
          tc_fence_after();
          uint32_t r[16];
          fetch_acc16(r, KSP > 1);
          if (row_valid) {
            float z[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) z[j] = __uint_as_float(r[j]) + s_bias[l * 16 + j];
            if (l == 0) {
              const unsigned ew[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                z[2 * j] += __uint_as_float(ew[j] << 16);
                z[2 * j + 1] += __uint_as_float(ew[j] & 0xffff0000u);
              }
            }
            __nv_bfloat16 hq[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float cn, hn;
              lstm_gates(z[4 * u], z[4 * u + 1], z[4 * u + 2], z[4 * u + 3], c_state[l][u], cn, hn);
              c_state[l][u] = cn;
              hq[u] = __float2bfloat16_rn(hn);
            }
            const uint2 pk = *reinterpret_cast<const uint2*>(hq);
            __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(p.xbuf[l]) + (size_t)(par ^ 1) * B * K +
                                (size_t)row * K + (K - Ud) + slice * 4;
            *reinterpret_cast<uint2*>(xs) = pk;
            if (l + 1 < L) {
              __nv_bfloat16* xu = reinterpret_cast<__nv_bfloat16*>(p.xbuf[l + 1]) + (size_t)par * B * (2 * Ud) +
                                  (size_t)row * (2 * Ud) + slice * 4;
              *reinterpret_cast<uint2*>(xu) = pk;
            }
          }
          tc_fence_before();
          if (fine && l == 0 && tid == 128) fine_stamp(3);
        }
        acc_parity ^= 1u;
      }
      dt_grid_barrier(p.bar, epoch);
      stamp(1 + l);
    }
    // ---------------- query layer (bahdanau): q = h_top . W_q ----------------
    if (bahdanau) {
      if (q_cta) {
        if (!q_ksplit) gemm_phase(&p.tmQ[par ^ 1], base + (uint32_t)p.off_wq, 0, Ud / 64, 16, IDESC, false);
        else gemm_phase(&p.tmQ[par ^ 1], base + (uint32_t)p.off_wq, crank * (Ud / 64 / KSP), Ud / 64 / KSP, 64, IDESC64, false);
        if (row_thread) {
          if (q_ksplit && tid == 128) mbar_expect_tx(pbar, part_bytes);
          mbar_wait(tfull, acc_parity);
          tc_fence_after();
          uint32_t r[16];
          fetch_acc16(r, q_ksplit);
          if (row_valid) {
            float4* qd = reinterpret_cast<float4*>(p.qbuf + (size_t)row * Ud + slice * 16);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              qd[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                  __uint_as_float(r[4 * j + 3]));
          }
          tc_fence_before();
        }
        acc_parity ^= 1u;
      }
      dt_grid_barrier(p.bar, epoch);
      stamp(5);
    }
    // ---------------- phase B: attention + context + logits ----------------
    {
      const int Dh = D / 2;
      const int Ktop = Kl[L - 1];
      const __nv_bfloat16* Htop = reinterpret_cast<const __nv_bfloat16*>(p.xbuf[L - 1]) + (size_t)(par ^ 1) * B * Ktop + (Ktop - Ud);
      for (int item = blockIdx.x; item < 2 * B; item += gridDim.x) {
        const int b = item >> 1, half = item & 1;
        const int len = min(d.mem_len[b], Tm);
        const __nv_bfloat16* keys = reinterpret_cast<const __nv_bfloat16*>(d.keys) + (size_t)b * Tm * Ud;
        const __nv_bfloat16* vals = reinterpret_cast<const __nv_bfloat16*>(d.values) + (size_t)b * Tm * D + (size_t)half * Dh;
        if (warp == 8) continue;  // the TMA producer warp has no role in this phase
        // ---- consumers: warps 0..7 ----
        for (int u = tid; u < Ud; u += 256)
          s_q[u] = bahdanau ? __ldcg(p.qbuf + (size_t)b * Ud + u) : __bfloat162float(Htop[(size_t)b * Ktop + u]);
        consumer_sync();
        // scores: warp per memory position; lane owns the 8-element chunks c8 = lane, lane+32 of the
        // depth (conflict-free 16-byte LDS of the staged keys); its slice of the query (and of
        // attention_v) lives in registers for the whole item when Ud <= 512
        // In cluster mode the two CTAs of an utterance (consecutive ranks of one cluster) split the depth of the
        // score reduction: each reads half of every key row and evaluates half of the tanh, then they swap partial
        // scores through DSMEM.  Otherwise every CTA scores the full depth.
        const bool pair_split = KSP > 1 && !p.no_pair_split;
        const int n_c8 = Ud / 8;
        const int c8_lo = pair_split ? half * (n_c8 / 2) : 0;
        const int c8_hi = pair_split ? c8_lo + n_c8 / 2 : n_c8;
        const bool reg_path = (c8_hi - c8_lo) <= 64;
        float qreg[16], vreg[16];
        if (reg_path) {
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int c8 = c8_lo + lane + 32 * cc;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              qreg[cc * 8 + j] = (c8 < c8_hi) ? s_q[c8 * 8 + j] : 0.f;
              vreg[cc * 8 + j] = (c8 < c8_hi && bahdanau) ? s_v[c8 * 8 + j] : 0.f;
            }
          }
        }
        // PV[b] (f32 [len][V]) is prefetched into the idle TMA ring with cp.async while scores and context run
        const bool pv_smem = half == 0 && (V % 4 == 0) && (size_t)Tm * V * 4 <= (size_t)NST * DT_STAGE;
        if (pv_smem) {
          const float* pvb = d.pv + (size_t)b * Tm * d.pv_ld;
          const int v4 = V / 4;
          for (int i = tid; i < len * v4; i += 256) {
            const int row = i / v4, c = i - row * v4;
            const uint32_t dst = ring + (uint32_t)((row * V + 4 * c) * 4);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(pvb + (size_t)row * d.pv_ld + 4 * c) : "memory");
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
        }
        // keys stream straight from L2 into registers (a cp.async.bulk ring was measured at ~1.7 us per stage of
        // latency; 4 rows x 2 chunks of 16 bytes in flight per lane does better): warp w takes rows w, w+8, ...
        auto load_keys = [&](int r0, uint4 (*kk)[2]) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = r0 + 8 * i;
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
              const int c8 = c8_lo + lane + 32 * cc;
              kk[i][cc] = make_uint4(0u, 0u, 0u, 0u);
              if (reg_path && r < len && c8 < c8_hi) kk[i][cc] = __ldg(reinterpret_cast<const uint4*>(keys + (size_t)r * Ud) + c8);
            }
          }
        };
        auto score_rows = [&](int r0, uint4 (*kk)[2]) {
          if (reg_path) {
            // four rows at a time with independent accumulators: the MUFU.TANH -> FFMA chains of different rows
            // overlap (one row alone is latency-bound on the special-function unit)
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
              if (c8_lo + 32 * cc >= c8_hi) break;  // warp-uniform: this half-depth has one chunk per lane
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float k0[4], k1[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const unsigned kw = j == 0 ? kk[i][cc].x : (j == 1 ? kk[i][cc].y : (j == 2 ? kk[i][cc].z : kk[i][cc].w));
                  k0[i] = __uint_as_float(kw << 16);
                  k1[i] = __uint_as_float(kw & 0xffff0000u);
                }
                if (bahdanau) {
                  float t0[4], t1[4];
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    t0[i] = tanh_mufu(k0[i] + qreg[cc * 8 + 2 * j]);
                    t1[i] = tanh_mufu(k1[i] + qreg[cc * 8 + 2 * j + 1]);
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
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = r0 + 8 * i;
              const float sacc = warp_sum(acc[i]);
              if (lane == 0 && r < len) s_score[r] = sacc;
            }
            return;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = r0 + 8 * i;
            if (r >= len) break;
            float acc = 0.f;
            const uint4* kr = reinterpret_cast<const uint4*>(keys + (size_t)r * Ud);
            for (int c8 = c8_lo + lane; c8 < c8_hi; c8 += 32) {
              const uint4 k4 = __ldg(kr + c8);
              const unsigned kw[4] = {k4.x, k4.y, k4.z, k4.w};
              const int u0 = c8 * 8;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float k0 = __uint_as_float(kw[j] << 16), k1 = __uint_as_float(kw[j] & 0xffff0000u);
                if (bahdanau) {
                  acc = fmaf(s_v[u0 + 2 * j], tanh_mufu(k0 + s_q[u0 + 2 * j]), acc);
                  acc = fmaf(s_v[u0 + 2 * j + 1], tanh_mufu(k1 + s_q[u0 + 2 * j + 1]), acc);
                } else {
                  acc = fmaf(k0, s_q[u0 + 2 * j], acc);
                  acc = fmaf(k1, s_q[u0 + 2 * j + 1], acc);
                }
              }
            }
            acc = warp_sum(acc);
            if (lane == 0) s_score[r] = acc;
          }
        };
        // keys stream straight from L2 into registers, software-pipelined: while the warp scores 4 rows the next
        // 4 rows (2 x 16 bytes per lane each) are already in flight.  Warp w takes rows w, w+8, ...
        {
          uint4 ka[4][2], kb2[4][2];
          load_keys(warp, ka);
          for (int r0 = warp; r0 < len; r0 += 64) {
            load_keys(r0 + 32, kb2);
            score_rows(r0, ka);
            load_keys(r0 + 64, ka);
            score_rows(r0 + 32, kb2);
          }
        }
        consumer_sync();
        if (pair_split) {
          float* peer_buf = s_peer + (size_t)(items_done & 1) * p.tm_pad;
          const uint32_t dst0 = mapa_cl(smem_u32(peer_buf), (uint32_t)(crank ^ 1));
          const uint32_t rb = mapa_cl(sbar, (uint32_t)(crank ^ 1));
          if (tid == 0) mbar_expect_tx(sbar, (uint32_t)(p.tm_pad * 4));
          for (int i = tid; i < p.tm_pad / 4; i += 256) {
            const float4 v = *reinterpret_cast<const float4*>(s_score + 4 * i);
            st_async_v4_cl(dst0 + 16 * i, make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w)), rb);
          }
          consumer_sync();  // every thread has read its chunk of the partial scores before they are overwritten below
          mbar_wait(sbar, sparity);
          sparity ^= 1u;
          for (int tm = tid; tm < len; tm += 256) {  // same order in both CTAs: (depth half 0) + (depth half 1)
            const float own = s_score[tm], oth = peer_buf[tm];
            s_score[tm] = half == 0 ? own + oth : oth + own;
          }
          ++items_done;
          consumer_sync();
        }
        fineB(0);
        if (monotonic) {
          // tf.contrib.seq2seq.monotonic_attention(mode='parallel'): two prefix sums along memory time
          float* sa = s_scan;
          float* sb = s_scan + p.tm_pad;
          const float tiny = 1.17549435e-38f;
          const float* prev = p.align_state + (size_t)b * Tm;
          for (int tm = tid; tm < Tm; tm += 256) {
            const float pc = (tm < len) ? sigmoidf_acc(s_score[tm] + d.score_bias) : 0.f;
            s_score[tm] = pc;
            sa[tm] = logf(fminf(fmaxf(1.f - pc, tiny), 1.f));
          }
          consumer_sync();
          for (int off = 1; off < Tm; off <<= 1) {  // inclusive Hillis-Steele scan of the logs
            for (int tm = tid; tm < Tm; tm += 256) sb[tm] = sa[tm] + (tm >= off ? sa[tm - off] : 0.f);
            consumer_sync();
            float* tmp = sa; sa = sb; sb = tmp;
          }
          for (int tm = tid; tm < Tm; tm += 256) {
            const float cpv = expf(tm > 0 ? sa[tm - 1] : 0.f);  // exclusive cumulative product of (1 - p)
            const float pv = (t == 0) ? (tm == 0 ? 1.f : 0.f) : __ldcg(prev + tm);
            sb[tm] = pv / fminf(fmaxf(cpv, 1e-10f), 1.f);
            s_score[tm] = s_score[tm] * cpv;  // p * cp
          }
          consumer_sync();
          float* ra = sb;
          float* rb = sa;
          for (int off = 1; off < Tm; off <<= 1) {
            for (int tm = tid; tm < Tm; tm += 256) rb[tm] = ra[tm] + (tm >= off ? ra[tm - off] : 0.f);
            consumer_sync();
            float* tmp = ra; ra = rb; rb = tmp;
          }
          for (int tm = tid; tm < Tm; tm += 256) s_score[tm] = s_score[tm] * ra[tm];
          consumer_sync();
          if (half == 0)
            for (int tm = tid; tm < Tm; tm += 256) __stcg(p.align_state + (size_t)b * Tm + tm, s_score[tm]);
        } else {
          float m = -INFINITY;
          for (int tm = tid; tm < len; tm += 256) m = fmaxf(m, s_score[tm]);
          m = warp_max(m);
          if (lane == 0) s_red[warp] = m;
          consumer_sync();
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
          consumer_sync();
          sum = 0.f;
#pragma unroll
          for (int w = 0; w < 8; ++w) sum += s_red[8 + w];
          for (int tm = tid; tm < Tm; tm += 256) s_score[tm] = s_score[tm] / sum;
          consumer_sync();
        }
        fineB(1);
        if (d.alignment && half == 0) {
          float* ar = d.alignment + ((size_t)b * d.max_steps + t) * Tm;
          for (int tm = tid; tm < Tm; tm += 256) ar[tm] = s_score[tm];
        }
        // context over this item's channel half: thread = (8-channel group, row parity)
        {
          const int cg = tid & 127, pr = tid >> 7;
          const bool has = cg < Dh / 8;
          float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (has) {
            const uint4* vp = reinterpret_cast<const uint4*>(vals) + cg;  // row stride D/8 uint4
            const size_t vstride = (size_t)(D / 8);
            auto load_vals = [&](int r0, uint4* vv) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int r = r0 + 2 * i;
                vv[i] = make_uint4(0u, 0u, 0u, 0u);
                if (r < len) vv[i] = __ldg(vp + (size_t)r * vstride);
              }
            };
            auto accum = [&](int r0, const uint4* vv) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int r = r0 + 2 * i;
                const float a = (r < len) ? s_score[r] : 0.f;
                const unsigned w4[4] = {vv[i].x, vv[i].y, vv[i].z, vv[i].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  acc[2 * e] = fmaf(a, __uint_as_float(w4[e] << 16), acc[2 * e]);
                  acc[2 * e + 1] = fmaf(a, __uint_as_float(w4[e] & 0xffff0000u), acc[2 * e + 1]);
                }
              }
            };
            // rows pr, pr+2, ...: 8 x 16 bytes in flight per thread while the previous 8 rows are accumulated
            uint4 va[8], vb2[8];
            load_vals(pr, va);
            for (int r0 = pr; r0 < len; r0 += 32) {
              load_vals(r0 + 16, vb2);
              accum(r0, va);
              load_vals(r0 + 32, va);
              accum(r0 + 16, vb2);
            }
          }
          if (has && pr == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) s_comb[cg * 8 + i] = acc[i];
          }
          consumer_sync();
          if (has && pr == 0) {
            __nv_bfloat16 o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = __float2bfloat16_rn(acc[i] + s_comb[cg * 8 + i]);
            __nv_bfloat16* att = reinterpret_cast<__nv_bfloat16*>(p.xbuf[0]) + (size_t)(par ^ 1) * B * (D + Ud) +
                                 (size_t)b * (D + Ud) + (size_t)half * Dh + cg * 8;
            *reinterpret_cast<uint4*>(att) = *reinterpret_cast<const uint4*>(o);
          }
        }
        fineB(2);
        // logits = a . PV[b] + bias and their argmax (lowest index wins ties), by the half-0 item of the utterance
        if (half == 0) {
          const int vi = tid & 63, g = tid >> 6;
          const float* pvb = d.pv + (size_t)b * Tm * d.pv_ld;
          const float* pvs = reinterpret_cast<const float*>(ring_ptr);
          if (pv_smem) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            consumer_sync();
          }
          float best = -INFINITY;
          int bi = 0x7fffffff;
          for (int vb = 0; vb < V; vb += 64) {
            const int v = vb + vi;
            float acc = 0.f;
            if (v < V) {
              if (pv_smem) {
#pragma unroll 8
                for (int tm = g; tm < len; tm += 4) acc = fmaf(s_score[tm], pvs[tm * V + v], acc);
              } else {
                const float* pc = pvb + v;
#pragma unroll 16
                for (int tm = g; tm < len; tm += 4) acc = fmaf(s_score[tm], __ldg(pc + (size_t)tm * d.pv_ld), acc);
              }
            }
            s_lp[g * 64 + vi] = acc;
            consumer_sync();
            if (tid < 64 && vb + tid < V) {
              const float sv = d.b_proj[vb + tid] + ((s_lp[tid] + s_lp[64 + tid]) + (s_lp[128 + tid] + s_lp[192 + tid]));
              d.logits[((size_t)b * d.max_steps + t) * V + vb + tid] = sv;
              if (sv > best) { best = sv; bi = vb + tid; }
            }
            consumer_sync();
          }
          if (tid < 64) {  // two warps: shuffle argmax inside each, then combine through shared memory
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const float ob = __shfl_xor_sync(0xffffffffu, best, o);
              const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
              if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (lane == 0) { s_red[16 + 2 * warp] = best; s_red[17 + 2 * warp] = __int_as_float(bi); }
          }
          consumer_sync();
          if (tid == 0) {
            float b0 = s_red[16], b1 = s_red[18];
            int i0 = __float_as_int(s_red[17]), i1 = __float_as_int(s_red[19]);
            if (b1 > b0 || (b1 == b0 && i1 < i0)) { b0 = b1; i0 = i1; }
            __stcg(p.next_ids + b, i0 == 0x7fffffff ? 0 : i0);
          }
        }
      }
      fineB(3);
      stamp(7);  // own attention work done, before the barrier
      dt_grid_barrier(p.bar, epoch);
      stamp(6);
    }
    ++t;
  }
  if (timing)
    for (int i = 0; i < 8; ++i) p.dbg[i] = tacc[i];
  if (fine) {  // written by the role threads that took the stamps (others hold zeros)
    if (warp == 0 && tfine[0]) { p.dbg[8] = tfine[0]; p.dbg[9] = tfine[1]; }
    if (tid == 128) { p.dbg[10] = tfine[2]; p.dbg[11] = tfine[3]; }
    if (tid == 0) for (int i = 0; i < 4; ++i) p.dbg[12 + i] = tfineB[i];
  }
  if (blockIdx.x == 0 && tid == 0) *d.n_steps = t;
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem_base) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct DecTcPlan {
  bool ok;
  int n_stages, n_stages_a, stage_a, tm_pad, ksplit, off_part;
  int off_w[4], off_wq, off_ring, off_misc;
  size_t smem;
  size_t ws_off_x[4], ws_off_q, ws_off_align, ws_off_ids, ws_off_bar, ws_off_dbg, ws_total;
};

static DecTcPlan dec_tc_plan(const plas_dec_desc& d) {
  DecTcPlan pl;
  memset(&pl, 0, sizeof(pl));
  const bool shape_ok = d.dtype == PLAS_BF16 && d.w_cell_tc[0] && d.pv && d.B <= 128 && d.Ud % 64 == 0 &&
                        d.D % 64 == 0 && d.D <= 2048 && d.Ud <= 2048 && (d.Ud / 4) <= num_sms() && d.Tm <= 4096 &&
                        (d.attention_type != PLAS_ATT_BAHDANAU || d.w_query_tc);
  if (!shape_ok) return pl;
  int off = 0;
  for (int l = 0; l < d.n_layers; ++l) {
    const int K = (l == 0) ? d.D + d.Ud : 2 * d.Ud;
    pl.off_w[l] = off;
    off += (K / 64) * 2048;
  }
  pl.off_wq = off;
  if (d.attention_type == PLAS_ATT_BAHDANAU) off += (d.Ud / 64) * 2048;
  pl.off_ring = off;  // multiples of 2048 -> 1024-aligned
  pl.tm_pad = (d.Tm + 3) & ~3;
  const int misc = 1536 + 4 * (64 + 64 + 256 + 2 * d.Ud + 5 * pl.tm_pad + d.D / 2) + 64;
  const int avail = 227 * 1024 - 1024 - off - misc;
  int ns = avail / DT_STAGE;
  if (ns > DT_MAX_STAGES) ns = DT_MAX_STAGES;
  if (ns < 2) return pl;
  pl.n_stages = ns;
  // LSTM phases: the TMA box holds min(B,128) rounded to 64 rows; with 64-row boxes the stage is 8 KB and
  // the M=128 MMA reads 8 KB past it (rows 64..127 = don't-care TMEM lanes), hence one stage of slack
  pl.stage_a = d.B <= 64 ? 8192 : 16384;
  pl.n_stages_a = d.B <= 64 ? 2 * ns - 1 : ns;
  if (pl.n_stages_a > DT_MAX_STAGES) pl.n_stages_a = DT_MAX_STAGES;
  // K-split mode (clusters of 4): the tail of the ring holds the peers' partial sums during the LSTM phases
  // ([4 senders][64 or 128 rows][16] f32)
  const int part_bytes = d.B <= 64 ? 16384 : 32768;
  pl.ksplit = 1;
  pl.off_part = off + ns * DT_STAGE - part_bytes;
  {
    bool ok = (d.Ud / 4) % 4 == 0 && ns * DT_STAGE - part_bytes - (d.B <= 64 ? 8192 : 0) >= 2 * pl.stage_a;
    for (int l = 0; l < d.n_layers; ++l) {
      const int K = (l == 0) ? d.D + d.Ud : 2 * d.Ud;
      ok = ok && (K / 64) % 4 == 0;
    }
    const char* e = getenv("PLAS_DEC_KSPLIT");
    if (e && atoi(e) == 1) ok = false;
    if (ok) {
      pl.ksplit = 4;
      pl.n_stages_a = (ns * DT_STAGE - part_bytes - (d.B <= 64 ? 8192 : 0)) / pl.stage_a;
      if (pl.n_stages_a > DT_MAX_STAGES) pl.n_stages_a = DT_MAX_STAGES;
    }
  }
  pl.off_misc = off + ns * DT_STAGE;
  pl.smem = (size_t)pl.off_misc + misc + 1024;
  size_t w = 0;
  auto take = [&](size_t bytes) { size_t o = w; w += (bytes + 255) & ~size_t(255); return o; };
  for (int l = 0; l < 4; ++l) {
    const size_t K = (l == 0) ? (size_t)d.D + d.Ud : (size_t)2 * d.Ud;
    pl.ws_off_x[l] = take(l < d.n_layers ? 2 * (size_t)d.B * K * 2 : 0);
  }
  pl.ws_off_q = take((size_t)d.B * d.Ud * 4);
  pl.ws_off_align = take((size_t)d.B * d.Tm * 4);
  pl.ws_off_ids = take((size_t)d.B * 4);
  pl.ws_off_bar = take(4);
  pl.ws_off_dbg = take(128);
  pl.ws_total = w;
  pl.ok = true;
  return pl;
}

size_t dec_tc_workspace_bytes(const plas_dec_desc& d) {
  const DecTcPlan pl = dec_tc_plan(d);
  return pl.ok ? pl.ws_total : 0;
}

// Returns PLAS_OK after a launch, 1 when the shape is not eligible (caller falls back), <0 on error.
int dec_tc_launch(const plas_dec_desc& d, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  const char* force = getenv("PLAS_DEC_IMPL");
  if (force && strcmp(force, "simt") == 0) return 1;
  const DecTcPlan pl = dec_tc_plan(d);
  if (!pl.ok) return 1;
  PLAS_REQUIRE(workspace_bytes >= pl.ws_total, "decoder(tc): workspace %zu < %zu", workspace_bytes, pl.ws_total);
  PLAS_CUDA(cudaMemsetAsync(workspace, 0, pl.ws_total, stream));
  PLAS_CUDA(cudaMemsetAsync(d.seq_len, 0, (size_t)d.B * 4, stream));
  PLAS_CUDA(cudaMemsetAsync(d.n_steps, 0, 4, stream));
  if (d.max_steps == 0) return PLAS_OK;

  DecTcArgs a;
  memset(&a, 0, sizeof(a));
  a.d = d;
  unsigned char* ws = (unsigned char*)workspace;
  for (int l = 0; l < 4; ++l) {
    a.xbuf[l] = ws + pl.ws_off_x[l];
    a.w_tc[l] = (const unsigned char*)d.w_cell_tc[l];
    a.off_w[l] = pl.off_w[l];
  }
  a.wq_tc = (const unsigned char*)d.w_query_tc;
  a.qbuf = (float*)(ws + pl.ws_off_q);
  a.align_state = (float*)(ws + pl.ws_off_align);
  a.next_ids = (int*)(ws + pl.ws_off_ids);
  a.bar = (unsigned*)(ws + pl.ws_off_bar);
  a.dbg = getenv("PLAS_DEBUG") ? (unsigned long long*)(ws + pl.ws_off_dbg) : nullptr;
  a.n_stages = pl.n_stages;
  a.n_stages_a = pl.n_stages_a;
  a.stage_a = pl.stage_a;
  a.ksplit = pl.ksplit;
  a.no_pair_split = getenv("PLAS_DEC_NOPAIR") ? 1 : 0;
  a.off_part = pl.off_part;
  a.off_wq = pl.off_wq;
  a.off_ring = pl.off_ring;
  a.off_misc = pl.off_misc;
  a.tm_pad = pl.tm_pad;
  for (int l = 0; l < d.n_layers; ++l) {
    const long long K = (l == 0) ? (long long)d.D + d.Ud : 2LL * d.Ud;
    for (int par = 0; par < 2; ++par) {
      int rc = make_map_bf16(&a.tmX[l][par], a.xbuf[l] + (size_t)par * d.B * K * 2, d.B, (int)K, K, pl.stage_a / 128);
      if (rc) return rc;
    }
  }
  {
    const int l = d.n_layers - 1;
    const long long K = (l == 0) ? (long long)d.D + d.Ud : 2LL * d.Ud;
    for (int par = 0; par < 2; ++par) {
      int rc = make_map_bf16(&a.tmQ[par], a.xbuf[l] + ((size_t)par * d.B * K + (size_t)(K - d.Ud)) * 2, d.B, d.Ud, K,
                             pl.stage_a / 128);
      if (rc) return rc;
    }
  }
  PLAS_CUDA(cudaFuncSetAttribute(decoder_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  int per_sm = 0;
  PLAS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decoder_tc_kernel, DT_THREADS, pl.smem));
  PLAS_REQUIRE(per_sm >= 1, "decoder(tc): kernel does not fit on an SM (%zu bytes of shared memory)", pl.smem);
  int grid = num_sms();
  const int want = (d.Ud / 4 > 2 * d.B) ? d.Ud / 4 : 2 * d.B;
  if (grid > want) grid = want;
  // K-split mode: clusters of 4 consecutive CTAs (plain cluster launch -- all clusters must be co-resident for the
  // grid barrier, which the occupancy query checks; ncu cannot replay a cooperative cluster launch).  TMA multicast
  // of the activation boxes was tried instead and gave nothing: the phases are bound by the bytes DELIVERED to
  // each SM (~3.4 TB/s aggregate L2->SM), not by L2 reads, so only splitting K reduces them.
  const int nsl = d.Ud / 4;
  cudaError_t le = cudaErrorUnknown;
  if (a.ksplit > 1) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)nsl);
    cfg.blockDim = dim3(DT_THREADS);
    cfg.dynamicSmemBytes = pl.smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)a.ksplit;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int max_clusters = 0;
    cudaError_t qe = cudaOccupancyMaxActiveClusters(&max_clusters, decoder_tc_kernel, &cfg);
    if (qe == cudaSuccess && max_clusters * a.ksplit >= nsl) le = cudaLaunchKernelEx(&cfg, decoder_tc_kernel, a);
    if (getenv("PLAS_DEBUG"))
      fprintf(stderr, "[plas] decoder tc path: ksplit=%d grid=%d stages=%d (lstm ring %d x %d B) max_active_clusters=%d (%s) launch: %s\n",
              a.ksplit, nsl, pl.n_stages, pl.n_stages_a, pl.stage_a, max_clusters, cudaGetErrorString(qe), cudaGetErrorString(le));
    if (le != cudaSuccess) (void)cudaGetLastError();
  }
  if (le != cudaSuccess) {
    if (a.ksplit > 1) {  // fall back to the single-CTA K loop: needs the stage count of that mode
      a.ksplit = 1;
      a.n_stages_a = d.B <= 64 ? 2 * pl.n_stages - 1 : pl.n_stages;
      if (a.n_stages_a > DT_MAX_STAGES) a.n_stages_a = DT_MAX_STAGES;
    }
    int grid = num_sms();
    const int want = (d.Ud / 4 > 2 * d.B) ? d.Ud / 4 : 2 * d.B;
    if (grid > want) grid = want;
    if (getenv("PLAS_DEBUG"))
      fprintf(stderr, "[plas] decoder tc path: grid=%d stages=%d (lstm ring %d x %d B) smem=%zu\n", grid, pl.n_stages,
              a.n_stages_a, pl.stage_a, pl.smem);
    void* args[] = {(void*)&a};
    PLAS_CUDA(cudaLaunchCooperativeKernel((const void*)decoder_tc_kernel, dim3(grid), dim3(DT_THREADS), args, pl.smem, stream));
  }
  if (a.dbg) {  // debug only: synchronises
    unsigned long long h[16];
    PLAS_CUDA(cudaStreamSynchronize(stream));
    PLAS_CUDA(cudaMemcpy(h, a.dbg, sizeof(h), cudaMemcpyDeviceToHost));
    fprintf(stderr, "[plas] decoder tc phase time (us, CTA 0): prologue %.1f  lstm %.1f %.1f %.1f %.1f  query %.1f  attention %.1f (own work %.1f)\n",
            h[0] / 1e3, h[1] / 1e3, h[2] / 1e3, h[3] / 1e3, h[4] / 1e3, h[5] / 1e3, (h[6] + h[7]) / 1e3, h[7] / 1e3);
    fprintf(stderr, "[plas]   layer 0 detail (us after phase start, summed): first box %.1f  MMAs issued %.1f  accumulator ready %.1f  epilogue done %.1f\n",
            h[8] / 1e3, h[9] / 1e3, h[10] / 1e3, h[11] / 1e3);
    fprintf(stderr, "[plas]   attention detail (us after phase start, summed): scores %.1f  softmax %.1f  context %.1f  logits %.1f\n",
            h[12] / 1e3, h[13] / 1e3, h[14] / 1e3, h[15] / 1e3);
  }
  return PLAS_OK;
}

}  // namespace plas
