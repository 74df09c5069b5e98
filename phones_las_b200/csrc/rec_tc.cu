// K3 (tensor-core path, bf16) -- BiLSTM recurrence with W_hh stationary in TENSOR MEMORY.
//
// Same contract as rec.cu (tf.nn.bidirectional_dynamic_rnn over LSTMCell, reference
// las/ops.py:23-46, input projections hoisted into K2).  The step product is formulated transposed,
//     Z^T [128 gate columns x NR utterances] = W_slice [128 x U] . h_{s-1}^T [U x NR],
// so that one CTA's slice of W_hh (all four gates of 32 hidden units) is the M = 128 operand of
// tcgen05.mma and can live in TMEM for the whole sequence (A-from-TMEM "TS" form: 128 lanes x U/2
// packed-bf16 columns = 128 KB at U = 512), while the small, changing operand h_{s-1} (NR x U bf16,
// K-major SWIZZLE_128B) sits in shared memory.  A step costs U/16 TS-form MMAs of ~18 cycles each
// (measured, scripts/micro/mma_cost.cu: 10 + N/2; ~590 cycles at U = 512, NR = 16) instead of ~1100 cycles of mma.sync.
//
// The G = U/32 CTAs of one (direction, group of NR utterances) form a thread-block cluster; every
// CTA pushes its NR x 32 slice of h_s straight into the (swizzled) operand tile of all G CTAs with
// st.async, completing transaction bytes on each receiver's mbarrier -- h never leaves the chip and
// there is no cluster barrier, no atomics and no fence on the step path (protocol as in rec.cu).
//
// Epilogue: the TMEM lanes of a warp quadrant are ordered 8*gate + unit, so two tcgen05.ld.16x128b hand
// thread t all four gates of unit t/4 for one utterance (16 gate-math warps, one cell per thread and group) -- no
// transpose --, then the TF gate math (forget_bias 1.0, length masking, bw direction walking len-1-s) runs with the
// cell state in registers.  Groups hold <= 16 utterances (rows); the host spreads the batch over the clusters its
// budget allows (rec_tc_try) and pad rows are neither exchanged nor stored.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tcgen05.cuh"
#include "../../include/plas.h"

namespace plas {

constexpr int RT_GW = 16;                        // gate-math warps: 4 per TMEM lane quadrant, one (unit, utterance) cell per thread and group
constexpr int RT_THREADS2 = (RT_GW + 5) * 32;    // gate-math warps + 1 MMA-issuer warp + 4 publisher warps
constexpr int RT_NPUB = 128;      // publisher threads

struct RecTcArgs {
  plas_rec_desc d;
  const void* whh_tc;  // [ndir][G][128][U] bf16, row (TMEM lane) m = 32*q + 8*gate + u8 for unit 8*q + u8
  int n_groups;
  int rows;            // utterances per group (<= 16): groups are padded to the N = 16 of the MMA, pad rows are never exchanged
  unsigned long long* tdbg;  // optional [8] ns counters written by CTA 0 (PLAS_DEBUG)
};

__device__ __forceinline__ uint32_t rt_mapa(uint32_t local_saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void rt_st_async_v4(uint32_t raddr, const uint4& v, uint32_t rbar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(rbar)
               : "memory");
}
// named barriers between the live gate-math warps and the 128 publisher threads (double-buffered stage tile):
// ids 2,3 = "stage tile b is full", ids 4,5 = "stage tile b has been published"
__device__ __forceinline__ void rt_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void rt_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void rt_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
// 16 TMEM lanes x 8 columns in the mma-C-fragment distribution: thread t gets lane t/4 (r0,r1) and lane t/4+8
// (r2,r3), columns 2*(t%4) and 2*(t%4)+1
__device__ __forceinline__ void tmem_ld_16x256b(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_16x128b(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
}
// The gate math is MUFU-bound (16 transcendentals per clock and SM; 512 cells per group step): 7 instead of 10 MUFU
// instructions per cell by putting the three factors of the cell update, and the two of the output, over one common
// denominator -- c = (c_prev (1+b)(1+d) + n_j (1+a)) / ((1+a)(1+b)(1+d)) with a = e^-(zf+1), b = e^-zi, d = e^-2|zj| and
// n_j = tanh(zj) (1+d) (the Cephes polynomial below 0.625, 1-d above, as tanhf_fast) -- one reciprocal each.  The
// exponents of a, b, e are clamped to 2^40 (sigmoid error < 1e-12) so that the products stay finite.
__device__ __forceinline__ float rt_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rt_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rt_tanh_num(float x, float d) {  // tanh(x) * (1 + d), d = e^-2|x|
  const float x2 = x * x;
  float p = fmaf(x2, -5.70498872745e-3f, 2.06390887954e-2f);
  p = fmaf(p, x2, -5.37397155531e-2f);
  p = fmaf(p, x2, 1.33314422036e-1f);
  p = fmaf(p, x2, -3.33332819422e-1f);
  const float small = fmaf(x * x2, p, x) * (1.0f + d);
  return fabsf(x) < 0.625f ? small : copysignf(1.0f - d, x);
}
__device__ __forceinline__ void rt_lstm_gates(float zi, float zj, float zf, float zo, float c_prev, float& c, float& h) {
  constexpr float L2E = 1.4426950408889634f;
  const float a = rt_ex2(fminf((zf + 1.0f) * -L2E, 40.0f));
  const float b = rt_ex2(fminf(zi * -L2E, 40.0f));
  const float e = rt_ex2(fminf(zo * -L2E, 40.0f));
  const float d = rt_ex2(fabsf(zj) * (-2.0f * L2E));
  const float A = 1.0f + a, BD = (1.0f + b) * (1.0f + d);
  c = fmaf(c_prev, BD, rt_tanh_num(zj, d) * A) * rt_rcp(A * BD);
  const float g = rt_ex2(fabsf(c) * (-2.0f * L2E));
  h = rt_tanh_num(c, g) * rt_rcp((1.0f + e) * (1.0f + g));
}
// NG independent 16-utterance groups of one direction share a cluster (and the TMEM-resident weights) and are
// software-pipelined against each other: warp RT_GW waits for a group's h_{s-1} to land and issues its U/16 MMAs
// (asynchronous, own accumulator columns), while the gate-math warps run the gate math / exchange of the other
// group(s).  The exchange latency (~0.5 us) and the MMA time of one group hide behind the epilogue of the others.
template <int KS, int NG>
__global__ void __launch_bounds__(RT_THREADS2, 1) rec_tc_kernel(RecTcArgs p) {
  constexpr int U = KS * 16;
  constexpr int G = U / 32;          // CTAs per cluster
  constexpr int KB = U / 64;         // 64-wide k blocks of the h operand
  constexpr int NR = 16;             // utterances per group
  constexpr int WQ = RT_GW / 4;      // gate-math warps per TMEM lane quadrant
  constexpr int HR = NR / WQ;        // utterances per gate-math warp (8: two per thread, 4: one per thread)
  constexpr int PP = HR / 4;         // (unit, utterance) pairs per thread and group
  constexpr int TILE = NR * 128;     // bytes of one k block of the h operand
  constexpr int HBUF = KB * TILE;    // bytes of one h operand buffer
  constexpr int ACOL0 = 64;          // first TMEM column of the resident W slice; accumulator of group g: 16g..16g+15
  constexpr uint32_t TMEM_COLS = (ACOL0 + 8 * KS) <= 128 ? 128u : ((ACOL0 + 8 * KS) <= 256 ? 256u : 512u);
  constexpr uint32_t IDESC = umma_idesc_bf16(128, NR);
  extern __shared__ unsigned char rt_smem_raw[];
  const plas_rec_desc& d = p.d;
  const int B = d.B, T = d.T, ndir = d.ndir;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ci = blockIdx.x % G;     // rank inside the cluster
  const int cl = blockIdx.x / G;
  const int cpd = (p.n_groups + NG - 1) / NG;  // clusters per direction
  const int dir = cl / cpd;
  const int grp0 = (cl % cpd) * NG;  // first 16-utterance group of this cluster

  const uint32_t raw = smem_u32(rt_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* smem = rt_smem_raw + (base - raw);
  // [NG][2][HBUF] h operands (UMMA K-major SWIZZLE_128B) | [2][NR][32] bf16 stage tiles
  const uint32_t hbuf_u = base;
  __nv_bfloat16* s_stage = reinterpret_cast<__nv_bfloat16*>(smem + NG * 2 * HBUF);
  __shared__ int s_len[NG][NR];
  __shared__ int s_tmax[NG];
  __shared__ __align__(8) unsigned long long s_bar[NG][3];  // h buffers 0/1, MMA completion
  __shared__ uint32_t s_tmem;

  const int R = p.rows;              // real utterances per group
  if (tid < NG * NR) {
    const int gg = tid / NR, r = tid % NR;
    const int b = (grp0 + gg) * R + r;
    s_len[gg][r] = (grp0 + gg < p.n_groups && r < R && b < B) ? min(d.lengths[b], T) : 0;
  }
  __syncthreads();
  if (tid < NG) {
    int m = 0;
    for (int r = 0; r < NR; ++r) m = max(m, s_len[tid][r]);
    s_tmax[tid] = m;
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  int Tg[NG];
  int Tmax = 0;
#pragma unroll
  for (int gg = 0; gg < NG; ++gg) { Tg[gg] = s_tmax[gg]; Tmax = max(Tmax, Tg[gg]); }
  const uint32_t tmem_base = s_tmem;
  uint32_t hbar[NG][2], mbar[NG];
#pragma unroll
  for (int gg = 0; gg < NG; ++gg) {
    hbar[gg][0] = smem_u32(&s_bar[gg][0]);
    hbar[gg][1] = smem_u32(&s_bar[gg][1]);
    mbar[gg] = smem_u32(&s_bar[gg][2]);
  }
  // gate-math warps whose HR utterances are all padding (rows >= R) sit the sequence out: fewer MUFU instructions in flight, smaller barriers
  const int nbar = 4 * ((R + HR - 1) / HR) * 32 + RT_NPUB;
  const uint32_t step_bytes = (uint32_t)(G * R * 64);  // bytes every CTA receives per group step
  if (tid == 0) {
#pragma unroll
    for (int gg = 0; gg < NG; ++gg) {
      mbar_init(hbar[gg][0], 1);
      mbar_init(hbar[gg][1], 1);
      mbar_init(mbar[gg], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
    for (int gg = 0; gg < NG; ++gg) {
      if (Tg[gg] >= 2) mbar_expect_tx(hbar[gg][0], step_bytes);  // h_0
      if (Tg[gg] >= 3) mbar_expect_tx(hbar[gg][1], step_bytes);  // h_1
    }
  }

  // pad rows of a group (rows >= R) are never written by the exchange: keep them finite
  for (int i = tid; i < NG * 2 * HBUF / 16; i += RT_THREADS2) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

  // ---- resident W slice -> TMEM (lane m = gate column, 8*KS packed-bf16 columns) ---------------
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  if (warp < 4) {
    const uint4* wrow = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.whh_tc) +
                                                       (((size_t)dir * G + ci) * 128 + tid) * U);
#pragma unroll 4
    for (int c0 = 0; c0 < 8 * KS; c0 += 8) {  // 8 columns = 16 bf16 = two uint4
      const uint4 a = __ldg(wrow + c0 / 4), b = __ldg(wrow + c0 / 4 + 1);
      const uint32_t r[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      tmem_st8(tmem_base + lane_base + (uint32_t)(ACOL0 + c0), r);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  // every CTA of the cluster must be running, with its mbarriers initialised, before anyone pushes
  rt_cluster_sync();
  tc_fence_after();

  if (warp == RT_GW) {
    // ===== MMA issuer: per (step, group) wait for h_{s-1}, re-arm the buffer's barrier, issue, commit =====
    if (elect_one()) {
      unsigned long long t_wait = 0, t_issue = 0, t0 = 0, t1 = 0;
      const bool tm = p.tdbg != nullptr && blockIdx.x == 0;
      for (int s = 1; s < Tmax; ++s) {
        const int bsel = (s - 1) & 1;
#pragma unroll
        for (int gg = 0; gg < NG; ++gg) {
          if (s >= Tg[gg]) continue;
          if (tm) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
          mbar_wait(hbar[gg][bsel], (uint32_t)(((s - 1) >> 1) & 1));  // h_{s-1} of the whole group has landed
          if (tm) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); t_wait += t1 - t0; }
          if (s + 1 <= Tg[gg] - 2) mbar_expect_tx(hbar[gg][bsel], step_bytes);  // re-arm for h_{s+1}
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // st.async data -> tensor-core reads
          tc_fence_after();
          const uint32_t hb = hbuf_u + (uint32_t)((gg * 2 + bsel) * HBUF);
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            const uint64_t bdesc = umma_smem_desc(hb + (uint32_t)((ks >> 2) * TILE)) + (uint64_t)(2 * (ks & 3));
            umma_bf16_ts(tmem_base + (uint32_t)(gg * NR), tmem_base + (uint32_t)(ACOL0 + 8 * ks), bdesc, IDESC,
                         ks != 0 ? 1u : 0u);
          }
          umma_commit(mbar[gg]);
          if (tm) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0)); t_issue += t0 - t1; }
        }
      }
      if (tm) { p.tdbg[0] = t_wait; p.tdbg[1] = t_issue; }
    }
    __syncwarp();
  } else if (warp > RT_GW) {
    // ===== publisher warps RT_GW+1 .. RT_GW+4: push every staged 16 x 32 slice (a) into the swizzled h operand of every CTA of
    // the cluster -- thread = (destination, row, 16-byte chunk), so the four chunks of a row leave as one contiguous
    // 64-byte DSMEM segment (scattered 16-byte remote stores were measured 30% slower) -- and (b) for active rows to
    // the [B,T,ndir*U] layer output in HBM.  An SM sends only ~20 bytes/clk over the SM-to-SM network, so the
    // 16 KB of a group step keep the store unit busy for ~0.4 us: on their own warps these stores overlap the gate
    // math of the next group instead of stalling it. =====
    const int ptid = tid - (RT_GW + 1) * 32;
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(d.out);
    int n_item = 0;
    for (int s = 0; s < Tmax; ++s) {
#pragma unroll
      for (int gg = 0; gg < NG; ++gg) {
        if (s >= Tg[gg]) continue;
        const int sb = n_item & 1;
        ++n_item;
        const __nv_bfloat16* stg = s_stage + sb * (NR * 32);
        rt_bar_sync(2 + sb, nbar);  // wait until the gate-math warps have filled the tile
        if (s + 1 < Tg[gg]) {
          const uint32_t dst_buf = hbuf_u + (uint32_t)((gg * 2 + (s & 1)) * HBUF) + (uint32_t)((ci >> 1) * TILE);
          const uint32_t bar_l = hbar[gg][s & 1];
#pragma unroll
          for (int j = 0; j < (G * NR * 4 + RT_NPUB - 1) / RT_NPUB; ++j) {
            const int idx = ptid + j * RT_NPUB;
            if (idx < G * NR * 4) {
              const int rank = idx / (NR * 4), chunk = idx % (NR * 4);
              const int r = chunk >> 2, ch = chunk & 3;
              if (r < R) {  // pad rows of a group are never exchanged (their accumulator columns are ignored)
                const uint4 v = *reinterpret_cast<const uint4*>(stg + r * 32 + ch * 8);
                const int c = (ci & 1) * 4 + ch;  // 16-byte chunk inside the 128-byte row of the k block
                const uint32_t local = dst_buf + (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
                rt_st_async_v4(rt_mapa(local, (uint32_t)rank), v, rt_mapa(bar_l, (uint32_t)rank));
              }
            }
          }
        }
        if (ptid < NR * 4) {
          const int r = ptid >> 2, ch = ptid & 3;
          const int b = (grp0 + gg) * R + r;
          const int len = s_len[gg][r];
          if (r < R && b < B) {
            // active step: h at its own time index (bw walks len-1-s); past the length: the zero padding of frame s
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            int t = s;
            if (s < len) {
              v = *reinterpret_cast<const uint4*>(stg + r * 32 + ch * 8);
              t = dir ? (len - 1 - s) : s;
            }
            if (s < len || !d.out_zeroed) {
              __nv_bfloat16* odst = out + (size_t)b * d.out_batch_stride + (size_t)t * (ndir * U) + dir * U + ci * 32 + ch * 8;
              *reinterpret_cast<uint4*>(odst) = v;
            }
          }
        }
        rt_bar_arrive(4 + sb, nbar);  // tile published: the gate-math warps may overwrite it
      }
    }
    if (!d.out_zeroed) {
      // frames beyond the longest utterance of a group (and the even-length pad frame) are zero as well
      const int t_alloc = (int)(d.out_batch_stride / (ndir * U));
#pragma unroll
      for (int gg = 0; gg < NG; ++gg) {
        if (grp0 + gg >= p.n_groups) continue;
        const int n_t = t_alloc - Tg[gg];
        for (int i = ptid; i < n_t * NR * 4; i += RT_NPUB) {
          const int t = Tg[gg] + i / (NR * 4), r = (i / 4) % NR, ch = i & 3;
          const int b = (grp0 + gg) * R + r;
          if (r < R && b < B)
            *reinterpret_cast<uint4*>(out + (size_t)b * d.out_batch_stride + (size_t)t * (ndir * U) + dir * U + ci * 32 + ch * 8) =
                make_uint4(0u, 0u, 0u, 0u);
        }
      }
    }
  } else {
    // ===== gate-math warps 0 .. RT_GW-1 =====
    // gate-math mapping: lane -> unit u8 = lane/4 of this warp's 8 units, utterances 2*jq, 2*jq+1 of the warp half (w/4)
    const int u8 = lane >> 2, jq = lane & 3;
    const int unit = ci * 32 + (warp & 3) * 8 + u8;
    const int half0 = (warp >> 2) * HR;
    const __nv_bfloat16* xproj = reinterpret_cast<const __nv_bfloat16*>(d.xproj);
    const int NX = ndir * 4 * U;
    int len_p[NG][PP];
    const __nv_bfloat16* xnext[NG][PP];  // pre-activations of the step after the one held in xp (walks +-NX per step)
    const long long xstep = dir ? -(long long)NX : (long long)NX;
    float c_state[NG][PP], h_state[NG][PP];
    uint2 xp[NG][PP];
#pragma unroll
    for (int gg = 0; gg < NG; ++gg)
#pragma unroll
      for (int i = 0; i < PP; ++i) {
        const int r = half0 + PP * jq + i;
        len_p[gg][i] = s_len[gg][r];
        const __nv_bfloat16* xrow = xproj + ((size_t)min((grp0 + gg) * R + r, B - 1) * T) * NX + (size_t)dir * 4 * U + 4 * unit;
        c_state[gg][i] = 0.f;
        h_state[gg][i] = 0.f;
        xp[gg][i] = make_uint2(0u, 0u);
        const int t0 = dir ? (len_p[gg][i] - 1) : 0;
        xnext[gg][i] = xrow + (long long)max(t0, 0) * NX;
        if (0 < len_p[gg][i]) {
          xp[gg][i] = __ldg(reinterpret_cast<const uint2*>(xnext[gg][i]));
          xnext[gg][i] += xstep;
        }
      }
    uint32_t mma_parity[NG];
#pragma unroll
    for (int gg = 0; gg < NG; ++gg) mma_parity[gg] = 0;

    int n_item = 0;
    unsigned long long e_wait = 0, e_gate = 0, e_push = 0, e0 = 0, e1 = 0;
    const bool tme = p.tdbg != nullptr && blockIdx.x == 0 && tid == 0;
    for (int s = 0; s < (half0 < R ? Tmax : 0); ++s) {
#pragma unroll
      for (int gg = 0; gg < NG; ++gg) {
        if (s >= Tg[gg]) continue;  // uniform over the cluster
        if (tme) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(e0));
        const int sb = n_item & 1;
        __nv_bfloat16* stg = s_stage + sb * (NR * 32);
        if (n_item >= 2) rt_bar_sync(4 + sb, nbar);  // the publishers are done with this tile's previous content
        ++n_item;
        uint32_t zr[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};  // [i,j | gates 0,1 of rows 0,1] then [f,o]: see below
        if (s > 0) {
          mbar_wait(mbar[gg], mma_parity[gg]);
          if (tme) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(e1)); e_wait += e1 - e0; e0 = e1; }
          mma_parity[gg] ^= 1u;
          tc_fence_after();
          // accumulator: TMEM lane = gate column (8*gate + unit inside the warp's 32-lane quadrant), column =
          // utterance.  Two 16x256b loads hand thread t all four gates of unit t/4 for utterances 2*(t%4), +1 of
          // this warp half -- no shared-memory transpose.
          if (PP == 2) {
            tmem_ld_16x256b(tmem_base + lane_base + (uint32_t)(gg * NR + half0), zr);
            tmem_ld_16x256b(tmem_base + lane_base + (16u << 16) + (uint32_t)(gg * NR + half0), zr + 4);
          } else {  // 16x128b: thread t gets lane t/4 (r0) and lane t/4 + 8 (r1) of column t%4
            tmem_ld_16x128b(tmem_base + lane_base + (uint32_t)(gg * NR + half0), zr);
            tmem_ld_16x128b(tmem_base + lane_base + (16u << 16) + (uint32_t)(gg * NR + half0), zr + 2);
          }
          tmem_ld_wait();
          tc_fence_before();
        }
#pragma unroll
        for (int i = 0; i < PP; ++i) {
          const int rl = PP * jq + i;  // utterance inside the warp's HR
          // PP == 2: zr[0..1] = gate i (lane u8) rows 0,1; zr[2..3] = gate j (lane u8+8); zr[4..5] = gate f; zr[6..7] = gate o
          // PP == 1: zr[0..3] = gates i, j, f, o
          const float4 z = make_float4(__uint_as_float(zr[i]), __uint_as_float(zr[PP + i]), __uint_as_float(zr[2 * PP + i]),
                                       __uint_as_float(zr[3 * PP + i]));
          const __nv_bfloat162 x01 = *reinterpret_cast<const __nv_bfloat162*>(&xp[gg][i].x);
          const __nv_bfloat162 x23 = *reinterpret_cast<const __nv_bfloat162*>(&xp[gg][i].y);
          float cn, hn;
          rt_lstm_gates(z.x + __low2float(x01), z.y + __high2float(x01), z.z + __low2float(x23), z.w + __high2float(x23),
                          c_state[gg][i], cn, hn);
          const bool live = s < len_p[gg][i];
          c_state[gg][i] = live ? cn : c_state[gg][i];
          h_state[gg][i] = live ? bf16_round(hn) : h_state[gg][i];
          stg[(half0 + rl) * 32 + (warp & 3) * 8 + u8] = __float2bfloat16_rn(h_state[gg][i]);
          // prefetch this group's next gate pre-activations (consumed one full round later)
          xp[gg][i] = make_uint2(0u, 0u);
          if (s + 1 < len_p[gg][i]) {
            xp[gg][i] = __ldg(reinterpret_cast<const uint2*>(xnext[gg][i]));
            xnext[gg][i] += xstep;
          }
        }
        __threadfence_block();
        rt_bar_arrive(2 + sb, nbar);  // stage tile sb is full: the publisher warps take it from here
        if (tme) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(e1)); e_gate += e1 - e0; e0 = e1; }
        if (tme) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(e1)); e_push += e1 - e0; }
      }
    }
    if (tme) { p.tdbg[2] = e_wait; p.tdbg[3] = e_gate; p.tdbg[4] = e_push; }
#pragma unroll
    for (int gg = 0; gg < NG; ++gg)
#pragma unroll
      for (int i = 0; i < PP; ++i) {
        const int r = half0 + PP * jq + i;
        const int b = (grp0 + gg) * R + r;
        if (grp0 + gg < p.n_groups && r < R && b < B) {
          d.c_final[((size_t)dir * B + b) * U + unit] = c_state[gg][i];
          d.h_final[((size_t)dir * B + b) * U + unit] = h_state[gg][i];
        }
      }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <int KS, int NG>
static int rec_tc_try(RecTcArgs a, cudaStream_t stream, bool must_fit_one_wave, bool* launched) {
  const plas_rec_desc& d = a.d;
  constexpr int U = KS * 16, G = U / 32, NR = 16;
  *launched = false;
  // > 114 KB of shared memory: at most one CTA per SM, so the TMEM allocation can never contend
  size_t smem = 1024 + (size_t)NG * 2 * (U / 64) * NR * 128 + (size_t)2 * NR * 32 * 2;
  if (smem > 227 * 1024) return PLAS_OK;
  if (smem < 120 * 1024) smem = 120 * 1024;
  auto fn = rec_tc_kernel<KS, NG>;
  PLAS_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (G > 8) PLAS_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)G);
  cfg.blockDim = dim3(RT_THREADS2);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)G;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static int cached_max[1][64];  // the occupancy query costs ~20 us: cache it per (template, device)
  int dev = 0;
  PLAS_CUDA(cudaGetDevice(&dev));
  int max_clusters = dev < 64 ? cached_max[0][dev] : 0;
  cudaError_t qe = cudaSuccess;
  if (max_clusters == 0) {
    qe = cudaOccupancyMaxActiveClusters(&max_clusters, fn, &cfg);
    if (qe == cudaSuccess && dev < 64) cached_max[0][dev] = max_clusters;
  }
  if (qe != cudaSuccess || max_clusters < 1) {
    (void)cudaGetLastError();
    return PLAS_OK;
  }
  // The exchange is what bounds a step and its bytes per SM grow with the utterances a cluster carries, so the batch is
  // spread over as many clusters as the GPU can hold at once (7 of 16 CTAs on a B200): groups of `rows` <= 16
  // utterances, padded to the MMA's N = 16 (pad rows are neither exchanged nor stored).
  int budget = max_clusters;  // the caller may keep SMs free for concurrent work (plas_rec_desc.max_clusters)
  if (d.max_clusters > 0 && d.max_clusters < budget) budget = d.max_clusters;
  if (const char* fb = getenv("PLAS_REC_MAX_CLUSTERS")) budget = atoi(fb) > 0 && atoi(fb) < max_clusters ? atoi(fb) : max_clusters;
  const int cpd_fit = budget / d.ndir > 0 ? budget / d.ndir : 1;  // clusters per direction in one wave
  int rows = (d.B + cpd_fit * NG - 1) / (cpd_fit * NG);
  if (const char* fr = getenv("PLAS_REC_ROWS")) rows = atoi(fr);
  if (rows < 1) rows = 1;
  if (rows > NR) {
    if (must_fit_one_wave) return PLAS_OK;
    rows = NR;
  }
  a.rows = rows;
  a.n_groups = (d.B + rows - 1) / rows;
  const int clusters = d.ndir * ((a.n_groups + NG - 1) / NG);
  cfg.gridDim = dim3((unsigned)(clusters * G));
  if (getenv("PLAS_DEBUG"))
    fprintf(stderr, "[plas] rec tc path: U=%d NG=%d G=%d rows=%d groups=%d clusters=%d max_active_clusters=%d smem=%zu\n", U, NG,
            G, rows, a.n_groups, clusters, max_clusters, smem);
  unsigned long long* dbgbuf = nullptr;
  if (getenv("PLAS_DEBUG")) {
    PLAS_CUDA(cudaMalloc(&dbgbuf, 64));
    PLAS_CUDA(cudaMemset(dbgbuf, 0, 64));
  }
  a.tdbg = dbgbuf;
  PLAS_CUDA(cudaLaunchKernelEx(&cfg, fn, a));
  if (dbgbuf) {  // debug only: synchronises
    unsigned long long h[8];
    PLAS_CUDA(cudaStreamSynchronize(stream));
    PLAS_CUDA(cudaMemcpy(h, dbgbuf, 64, cudaMemcpyDeviceToHost));
    cudaFree(dbgbuf);
    fprintf(stderr, "[plas]   rec tc CTA 0 (us): issuer wait-h %.1f issue %.1f | epilogue wait-mma %.1f ld+gates %.1f publish %.1f  (T=%d)\n",
            h[0] / 1e3, h[1] / 1e3, h[2] / 1e3, h[3] / 1e3, h[4] / 1e3, d.T);
  }
  *launched = true;
  return PLAS_OK;
}

// Fewest groups per cluster such that the batch fits in one wave of clusters with <= 16 rows per group (measured at
// U = 512, us per step: B = 16: NG 1 x 6 rows 0.95, NG 2 x 3 rows 1.31 (round-2 start); B = 32: NG 1 x 11 rows 1.14; B = 64:
// NG 2 x 11 rows on 6 clusters 1.34, NG 2 x 16 rows on 4 clusters 1.47, NG 3 x 11 rows on 4 clusters 1.92; B = 128: NG 3 x 15 rows
// on 6 clusters 2.11, NG 4 x 11 rows 2.61, NG 4 x 16 rows on 4 clusters 2.94).
template <int KS>
static int rec_tc_launch_ks(const RecTcArgs& a, cudaStream_t stream) {
  bool launched = false;
  const char* force = getenv("PLAS_REC_NG");
  const int fng = force ? atoi(force) : 0;
  int rc;
  if (fng == 0 || fng == 1) {
    rc = rec_tc_try<KS, 1>(a, stream, fng == 0, &launched);
    if (rc || launched) return rc;
  }
  if (fng == 0 || fng == 2) {
    rc = rec_tc_try<KS, 2>(a, stream, fng == 0, &launched);
    if (rc || launched) return rc;
  }
  if (fng == 0 || fng == 3) {
    rc = rec_tc_try<KS, 3>(a, stream, fng == 0, &launched);
    if (rc || launched) return rc;
  }
  rc = rec_tc_try<KS, 4>(a, stream, false, &launched);
  if (rc || launched) return rc;
  return 1;
}

// Returns PLAS_OK after a launch, 1 when the shape is not eligible (caller falls back), <0 on error.
int rec_tc_launch(const plas_rec_desc& d, cudaStream_t stream) {
  const char* force = getenv("PLAS_REC_IMPL");
  if (force && strcmp(force, "tc") != 0) return 1;
  if (d.dtype != PLAS_BF16 || !d.whh_tc) return 1;
  RecTcArgs a;
  a.d = d;
  a.whh_tc = d.whh_tc;
  a.n_groups = 0;
  a.rows = 16;
  a.tdbg = nullptr;
  switch (d.U) {
    case 64: return rec_tc_launch_ks<4>(a, stream);
    case 128: return rec_tc_launch_ks<8>(a, stream);
    case 256: return rec_tc_launch_ks<16>(a, stream);
    case 512: return rec_tc_launch_ks<32>(a, stream);
    default: return 1;
  }
}

}  // namespace plas
