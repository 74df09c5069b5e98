// K3 -- persistent weight-stationary (bi)directional LSTM recurrence for sm_100a.
//
// Replaces tf.nn.bidirectional_dynamic_rnn over tf.nn.rnn_cell.LSTMCell (reference
// las/ops.py:23-46) once the input projections have been hoisted out of the time loop (K2).
//
// Decomposition.  The batch is cut into groups of 16 utterances; every (direction, group) pair is
// an independent recurrence and owns G = U / units_per_cta co-resident CTAs (cooperative launch).
// A CTA keeps its slice of W_hh -- all four gates of `upc` hidden units -- in shared memory for
// the whole sequence (bf16: pre-packed mma.m16n8k16 B fragments; f32: [k][4*upc] rows), keeps the
// cell state of its units in registers, and per time step
//   1. prefetches its gate pre-activations (K2 output) for step s,
//   2. waits on the group's step counter (L2 atomics, acquire/release), pulls h_{s-1} of the
//      whole group (16 x U) from the L2-resident exchange buffer with cp.async.cg,
//   3. computes z = xproj + h_{s-1} * W_hh on tensor cores (bf16) or FFMA (f32),
//   4. applies the TF gate math (order i,j,f,o; forget_bias 1.0) with length masking
//      (state frozen and output 0 for t >= len; the bw direction walks len-1-s),
//   5. publishes its h slice (exchange buffer + the [B,T,ndir*U] output) and bumps the counter.
// Only the CTAs of one (direction, group) ever synchronise with each other.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "../../include/plas.h"

namespace plas {

constexpr int REC_THREADS = 256;
constexpr int REC_ROWS = 16;  // utterances per group

struct RecArgs {
  plas_rec_desc d;
  void* hx;            // [2][ndir][Bpad][U] exchange buffer (dtype)
  unsigned* counters;  // [ndir][n_groups]
  int n_groups;        // total groups
  int group_offset;    // first group handled by this launch
  int groups_here;     // groups in this launch
  int upc;             // units per CTA
  int G;               // CTAs per (dir, group)
  int Bpad;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(s));
}
__device__ __forceinline__ void mma_bf16_16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void wait_counter(const unsigned* ctr, unsigned target) {
  unsigned spins = 0;
  while (ld_acquire_u32(ctr) < target) {
    if (++spins > (1u << 28)) __trap();  // a protocol bug must fail loudly, not hang the GPU
  }
}

// ------------------------------------------------------------------------------------------
// bf16 tensor-core variant: upc = 32 (8 warps x 4 units), 16 rows x 128 gate columns per CTA.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(REC_THREADS, 1) rec_bf16_kernel(RecArgs p) {
  extern __shared__ __align__(16) unsigned char rec_smem[];
  const plas_rec_desc& d = p.d;
  const int U = d.U, B = d.B, T = d.T, ndir = d.ndir;
  const int KS = U / 16;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  int bid = blockIdx.x;
  const int ci = bid % p.G; bid /= p.G;
  const int gi = p.group_offset + bid % p.groups_here;
  const int dir = bid / p.groups_here;
  const int row0 = gi * REC_ROWS;

  uint4* s_w = reinterpret_cast<uint4*>(rec_smem);                         // [8][KS][32] uint4
  const int hstride = U + 8;                                               // bf16 elements
  __nv_bfloat16* s_h = reinterpret_cast<__nv_bfloat16*>(rec_smem + (size_t)U * 256);   // [16][U+8]
  __nv_bfloat16* s_stage = s_h + REC_ROWS * hstride;                       // [16][32]
  __shared__ int s_len[REC_ROWS];
  __shared__ int s_tmax;

  {  // stage W_hh fragments once
    const uint4* src = reinterpret_cast<const uint4*>(d.whh) + ((size_t)dir * p.G + ci) * (size_t)(8 * KS * 32);
    for (int i = tid; i < 8 * KS * 32; i += REC_THREADS) s_w[i] = src[i];
  }
  if (tid < REC_ROWS) s_len[tid] = (row0 + tid < B) ? min(d.lengths[row0 + tid], T) : 0;
  __syncthreads();
  if (tid == 0) {
    int m = 0;
    for (int r = 0; r < REC_ROWS; ++r) m = max(m, s_len[r]);
    s_tmax = m;
  }
  __syncthreads();
  const int Tg = s_tmax;

  const __nv_bfloat16* xproj = reinterpret_cast<const __nv_bfloat16*>(d.xproj);
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(d.out);
  __nv_bfloat16* hx = reinterpret_cast<__nv_bfloat16*>(p.hx);
  unsigned* ctr = p.counters + dir * p.n_groups + gi;
  const int unit = ci * 32 + warp * 4 + q;  // this thread's hidden unit
  const int NX = ndir * 4 * U;              // xproj row length

  float c_state[2] = {0.f, 0.f};
  float h_state[2] = {0.f, 0.f};
  const int rows[2] = {g, g + 8};
  int len_r[2] = {s_len[g], s_len[g + 8]};

  for (int s = 0; s < Tg; ++s) {
    // 1. prefetch gate pre-activations (independent of h)
    uint2 xp[2];
    int t_idx[2];
    bool act[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      act[e] = s < len_r[e];
      t_idx[e] = dir ? (len_r[e] - 1 - s) : s;
      xp[e] = make_uint2(0u, 0u);
      if (act[e]) {
        const size_t off = ((size_t)(row0 + rows[e]) * T + t_idx[e]) * NX + (size_t)dir * 4 * U + 4 * unit;
        xp[e] = *reinterpret_cast<const uint2*>(xproj + off);
      }
    }
    float accA[4] = {0.f, 0.f, 0.f, 0.f}, accB[4] = {0.f, 0.f, 0.f, 0.f};
    if (s > 0) {
      // 2. wait for h_{s-1} of the whole group, pull it into shared memory
      if (tid == 0) wait_counter(ctr, (unsigned)(p.G * s));
      __syncthreads();
      const __nv_bfloat16* hsrc = hx + (((size_t)((s - 1) & 1) * ndir + dir) * p.Bpad + row0) * U;
      const int chunks_per_row = U / 8;
      for (int i = tid; i < REC_ROWS * chunks_per_row; i += REC_THREADS) {
        const int r = i / chunks_per_row, ch = i % chunks_per_row;
        cp_async16(s_h + r * hstride + ch * 8, hsrc + (size_t)r * U + ch * 8);
      }
      cp_async_wait_all();
      __syncthreads();
      // 3. z += h_{s-1} * W_hh   (A = h tile via ldmatrix, B = resident fragments)
      const __nv_bfloat16* arow = s_h + (lane & 15) * hstride + (lane >> 4) * 8;
      const uint4* wf = s_w + (size_t)warp * KS * 32 + lane;
#pragma unroll 4
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t a0, a1, a2, a3;
        ldmatrix_x4(a0, a1, a2, a3, arow + ks * 16);
        const uint4 w = wf[(size_t)ks * 32];
        mma_bf16_16816(accA, a0, a1, a2, a3, w.x, w.y);
        mma_bf16_16816(accB, a0, a1, a2, a3, w.z, w.w);
      }
    }
    // 4. gates (thread owns rows g and g+8 of unit `unit`)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const __nv_bfloat162 x01 = *reinterpret_cast<const __nv_bfloat162*>(&xp[e].x);
      const __nv_bfloat162 x23 = *reinterpret_cast<const __nv_bfloat162*>(&xp[e].y);
      const float zi = accA[2 * e] + __low2float(x01);
      const float zj = accA[2 * e + 1] + __high2float(x01);
      const float zf = accB[2 * e] + __low2float(x23);
      const float zo = accB[2 * e + 1] + __high2float(x23);
      float cn, hn;
      lstm_gates(zi, zj, zf, zo, c_state[e], cn, hn);
      if (act[e]) {
        c_state[e] = cn;
        h_state[e] = bf16_round(hn);
      }
      s_stage[rows[e] * 32 + warp * 4 + q] = __float2bfloat16_rn(h_state[e]);
    }
    __syncthreads();
    // 5. publish: 16 rows x 32 units = 64 chunks of 16 bytes
    if (tid < 64) {
      const int r = tid >> 2, ch = tid & 3;
      const int b = row0 + r;
      if (b < B) {
        const uint4 v = *reinterpret_cast<const uint4*>(s_stage + r * 32 + ch * 8);
        __nv_bfloat16* hdst = hx + (((size_t)(s & 1) * ndir + dir) * p.Bpad + b) * U + ci * 32 + ch * 8;
        *reinterpret_cast<uint4*>(hdst) = v;
        const int len = s_len[r];
        if (s < len) {
          const int t = dir ? (len - 1 - s) : s;
          __nv_bfloat16* odst = out + (size_t)b * d.out_batch_stride + (size_t)t * (ndir * U) + dir * U + ci * 32 + ch * 8;
          *reinterpret_cast<uint4*>(odst) = v;
        }
      }
    }
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      red_release_add_u32(ctr, 1u);
    }
  }
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int b = row0 + rows[e];
    if (b < B) {
      d.c_final[((size_t)dir * B + b) * U + unit] = c_state[e];
      d.h_final[((size_t)dir * B + b) * U + unit] = h_state[e];
    }
  }
}

// ------------------------------------------------------------------------------------------
// bf16 cluster variant (the fast path).
//
// * The G = U/32 CTAs that share one recurrence form ONE thread-block cluster (G <= 8 portable,
//   16 non-portable; U = 512 -> 16).
// * W_hh is weight-stationary in REGISTERS: thread (warp, lane) keeps the mma.m16n8k16 B fragments
//   of its warp's 16 gate columns for all U/16 k-steps (KS uint4 = 128 registers at U = 512), so the
//   only shared-memory traffic of a step is the h tile itself.
// * h_t never leaves the chip: every CTA pushes its 16 x 32 slice of h_t into the shared memory of
//   all G CTAs with st.async (a DSMEM store that completes transaction bytes on the RECEIVER's
//   mbarrier); a step is closed by each CTA's own mbarrier seeing G*1 KB -- no cluster barrier, no
//   L2 round trip, no atomics, no fence.  Buffer reuse needs no extra sync: a peer can only send
//   h_s after it received my h_{s-1}, which I send after my last read of the buffer h_s lands in.
// * NG = 2 interleaves two independent 16-utterance groups of the same direction in one cluster:
//   while group A's h is in flight the tensor cores work on group B, which hides the exchange
//   latency and lets 2*ndir*ceil(B/32) clusters cover the batch (only 7 clusters of 16 CTAs are
//   co-resident on a B200).
// * Gate pre-activations are prefetched into registers a full item ahead; the K loop runs on four
//   interleaved accumulator sets (summed 0+1+2+3) so the mma.sync dependency chain is U/64 deep and
//   the result does not depend on NG.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async_v4(uint32_t raddr, const uint4& v, uint32_t rbar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(rbar)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int KS, int NG>
__global__ void __launch_bounds__(REC_THREADS, 1) rec_bf16_cluster_kernel(RecArgs p) {
  constexpr int U = KS * 16;
  constexpr int G = U / 32;            // CTAs per cluster
  constexpr int R = REC_ROWS;          // utterances per group
  constexpr int HS = U + 8;            // padded row stride of the h tile (bf16 elements)
  extern __shared__ __align__(16) unsigned char rec_smem[];
  const plas_rec_desc& d = p.d;
  const int B = d.B, T = d.T, ndir = d.ndir;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int ci = blockIdx.x % G;       // rank inside the cluster (1-D clusters of G consecutive CTAs)
  const int cl = blockIdx.x / G;
  const int cpd = (p.n_groups + NG - 1) / NG;  // clusters per direction
  const int dir = cl / cpd;
  const int grp0 = (cl % cpd) * NG;    // first group of this cluster

  __nv_bfloat16* s_h = reinterpret_cast<__nv_bfloat16*>(rec_smem);            // [NG][2][R][HS]
  __nv_bfloat16* s_stage = s_h + (size_t)NG * 2 * R * HS;                     // [R][32]
  __shared__ int s_len[NG][R];
  __shared__ int s_tmax[NG];
  __shared__ __align__(8) unsigned long long s_bar[NG][2];

  // weight-stationary fragments: [dir][G][8 warps][KS][32 lanes] uint4 in global memory
  uint4 wreg[KS];
  {
    const uint4* src = reinterpret_cast<const uint4*>(d.whh) + (((size_t)dir * G + ci) * 8 + warp) * (size_t)(KS * 32) + lane;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) wreg[ks] = __ldg(src + (size_t)ks * 32);
  }
  if (tid < NG * R) {
    const int gg = tid / R, r = tid % R;
    const int b = (grp0 + gg) * R + r;
    s_len[gg][r] = (grp0 + gg < p.n_groups && b < B) ? min(d.lengths[b], T) : 0;
  }
  __syncthreads();
  if (tid < NG) {
    int m = 0;
    for (int r = 0; r < R; ++r) m = max(m, s_len[tid][r]);
    s_tmax[tid] = m;
  }
  __syncthreads();
  int Tg[NG];
  int Tmax = 0;
#pragma unroll
  for (int gg = 0; gg < NG; ++gg) { Tg[gg] = s_tmax[gg]; Tmax = max(Tmax, Tg[gg]); }

  const uint32_t step_bytes = (uint32_t)(G * R * 64);  // bytes every CTA receives per group step
  uint32_t bar_addr[NG][2];
#pragma unroll
  for (int gg = 0; gg < NG; ++gg) { bar_addr[gg][0] = smem_u32(&s_bar[gg][0]); bar_addr[gg][1] = smem_u32(&s_bar[gg][1]); }
  if (tid == 0) {
#pragma unroll
    for (int gg = 0; gg < NG; ++gg) {
      mbar_init(bar_addr[gg][0], 1);
      mbar_init(bar_addr[gg][1], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
    for (int gg = 0; gg < NG; ++gg) {
      if (Tg[gg] >= 2) mbar_expect_tx(bar_addr[gg][0], step_bytes);  // h_0
      if (Tg[gg] >= 3) mbar_expect_tx(bar_addr[gg][1], step_bytes);  // h_1
    }
  }

  const __nv_bfloat16* xproj = reinterpret_cast<const __nv_bfloat16*>(d.xproj);
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(d.out);
  const int unit = ci * 32 + warp * 4 + q;
  const int NX = ndir * 4 * U;
  int len_r[NG][2];
  const __nv_bfloat16* xrow[NG][2];
  float c_state[NG][2], h_state[NG][2];
  uint2 xp[NG][2];
#pragma unroll
  for (int gg = 0; gg < NG; ++gg)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int r = g + 8 * e;
      len_r[gg][e] = s_len[gg][r];
      xrow[gg][e] = xproj + ((size_t)min((grp0 + gg) * R + r, B - 1) * T) * NX + (size_t)dir * 4 * U + 4 * unit;
      c_state[gg][e] = 0.f;
      h_state[gg][e] = 0.f;
      xp[gg][e] = make_uint2(0u, 0u);
      if (0 < len_r[gg][e]) {
        const int t = dir ? (len_r[gg][e] - 1) : 0;
        xp[gg][e] = __ldg(reinterpret_cast<const uint2*>(xrow[gg][e] + (size_t)t * NX));
      }
    }
  const uint32_t s_h_addr = smem_u32(s_h);
  // every CTA of the cluster must be running, with its mbarriers initialised, before anyone pushes
  cluster_sync_all();

  for (int s = 0; s < Tmax; ++s) {
#pragma unroll
    for (int gg = 0; gg < NG; ++gg) {
      if (s >= Tg[gg]) continue;  // uniform over the cluster
      const int bsel = (s - 1) & 1;
      float acc[4][2][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int n = 0; n < 2; ++n)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[a][n][i] = 0.f;
      if (s > 0) {
        mbar_wait(bar_addr[gg][bsel], (uint32_t)(((s - 1) >> 1) & 1));  // h_{s-1} of the whole group has landed
        if (tid == 0 && s + 1 <= Tg[gg] - 2) mbar_expect_tx(bar_addr[gg][bsel], step_bytes);  // re-arm for h_{s+1}
        const __nv_bfloat16* hb = s_h + (size_t)(gg * 2 + bsel) * R * HS;
        const __nv_bfloat16* arow = hb + (lane & 15) * HS + (lane >> 4) * 8;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          uint32_t a0, a1, a2, a3;
          ldmatrix_x4(a0, a1, a2, a3, arow + ks * 16);
          mma_bf16_16816(acc[ks & 3][0], a0, a1, a2, a3, wreg[ks].x, wreg[ks].y);
          mma_bf16_16816(acc[ks & 3][1], a0, a1, a2, a3, wreg[ks].z, wreg[ks].w);
        }
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float z[2][2];  // [n-tile][pair element]: (i, j) and (f, o)
#pragma unroll
        for (int n = 0; n < 2; ++n)
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            float v = acc[0][n][2 * e + i];
#pragma unroll
            for (int a = 1; a < 4; ++a) v += acc[a][n][2 * e + i];
            z[n][i] = v;
          }
        const __nv_bfloat162 x01 = *reinterpret_cast<const __nv_bfloat162*>(&xp[gg][e].x);
        const __nv_bfloat162 x23 = *reinterpret_cast<const __nv_bfloat162*>(&xp[gg][e].y);
        const float zi = z[0][0] + __low2float(x01), zj = z[0][1] + __high2float(x01);
        const float zf = z[1][0] + __low2float(x23), zo = z[1][1] + __high2float(x23);
        float cn, hn;
        lstm_gates_fast(zi, zj, zf, zo, c_state[gg][e], cn, hn);
        if (s < len_r[gg][e]) {
          c_state[gg][e] = cn;
          h_state[gg][e] = bf16_round(hn);
        }
        s_stage[(g + 8 * e) * 32 + warp * 4 + q] = __float2bfloat16_rn(h_state[gg][e]);
        // prefetch this group's next gate pre-activations (consumed one full item later)
        xp[gg][e] = make_uint2(0u, 0u);
        if (s + 1 < len_r[gg][e]) {
          const int t = dir ? (len_r[gg][e] - 2 - s) : (s + 1);
          xp[gg][e] = __ldg(reinterpret_cast<const uint2*>(xrow[gg][e] + (size_t)t * NX));
        }
      }
      __syncthreads();
      // publish the staged 16 x 32 slice: (a) to every CTA of the cluster through DSMEM (not needed
      // after the group's last step), (b) for active rows to the [B,T,ndir*U] layer output in HBM
      if (s + 1 < Tg[gg]) {
        const uint32_t dst_buf = s_h_addr + (uint32_t)(((gg * 2 + (s & 1)) * R * HS) * 2);
#pragma unroll
        for (int j = 0; j < (G * 64 + REC_THREADS - 1) / REC_THREADS; ++j) {
          const int idx = tid + j * REC_THREADS;
          if (idx < G * 64) {
            const int rank = idx >> 6, chunk = idx & 63;
            const int r = chunk >> 2, ch = chunk & 3;
            const uint4 v = *reinterpret_cast<const uint4*>(s_stage + r * 32 + ch * 8);
            const uint32_t local = dst_buf + (uint32_t)((r * HS + ci * 32 + ch * 8) * 2);
            st_async_v4(mapa_u32(local, (uint32_t)rank), v, mapa_u32(bar_addr[gg][s & 1], (uint32_t)rank));
          }
        }
      }
      if (tid < 64) {
        const int r = tid >> 2, ch = tid & 3;
        const int b = (grp0 + gg) * R + r;
        const int len = s_len[gg][r];
        if (b < B && s < len) {
          const uint4 v = *reinterpret_cast<const uint4*>(s_stage + r * 32 + ch * 8);
          const int t = dir ? (len - 1 - s) : s;
          __nv_bfloat16* odst = out + (size_t)b * d.out_batch_stride + (size_t)t * (ndir * U) + dir * U + ci * 32 + ch * 8;
          *reinterpret_cast<uint4*>(odst) = v;
        }
      }
      __syncthreads();  // s_stage is rewritten by the next item's gate stage
    }
  }
#pragma unroll
  for (int gg = 0; gg < NG; ++gg)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int b = (grp0 + gg) * R + g + 8 * e;
      if (grp0 + gg < p.n_groups && b < B) {
        d.c_final[((size_t)dir * B + b) * U + unit] = c_state[gg][e];
        d.h_final[((size_t)dir * B + b) * U + unit] = h_state[gg][e];
      }
    }
}

// ------------------------------------------------------------------------------------------
// exact-fp32 SIMT variant (reference-precision mode).  upc units per CTA, thread = (unit, row group).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(REC_THREADS, 1) rec_f32_kernel(RecArgs p) {
  extern __shared__ __align__(16) unsigned char rec_smem[];
  const plas_rec_desc& d = p.d;
  const int U = d.U, B = d.B, T = d.T, ndir = d.ndir, upc = p.upc;
  const int tid = threadIdx.x;
  int bid = blockIdx.x;
  const int ci = bid % p.G; bid /= p.G;
  const int gi = p.group_offset + bid % p.groups_here;
  const int dir = bid / p.groups_here;
  const int row0 = gi * REC_ROWS;
  const int NC = 4 * upc;

  float* s_w = reinterpret_cast<float*>(rec_smem);  // [U][NC]
  float* s_h = s_w + (size_t)U * NC;                // [16][U]
  __shared__ int s_len[REC_ROWS];
  __shared__ int s_tmax;
  {
    const float* src = reinterpret_cast<const float*>(d.whh) + ((size_t)dir * p.G + ci) * (size_t)U * NC;
    for (int i = tid; i < U * NC; i += REC_THREADS) s_w[i] = src[i];
  }
  if (tid < REC_ROWS) s_len[tid] = (row0 + tid < B) ? min(d.lengths[row0 + tid], T) : 0;
  __syncthreads();
  if (tid == 0) {
    int m = 0;
    for (int r = 0; r < REC_ROWS; ++r) m = max(m, s_len[r]);
    s_tmax = m;
  }
  __syncthreads();
  const int Tg = s_tmax;

  const float* xproj = reinterpret_cast<const float*>(d.xproj);
  float* out = reinterpret_cast<float*>(d.out);
  float* hx = reinterpret_cast<float*>(p.hx);
  unsigned* ctr = p.counters + dir * p.n_groups + gi;
  const int NX = ndir * 4 * U;
  const int ul = tid % upc;             // local unit
  const int rg = tid / upc;             // row group
  const int nrg = REC_THREADS / upc;    // number of row groups
  const int unit = ci * upc + ul;
  constexpr int MAXR = 2;               // rows per thread (nrg >= 8 -> at most 2)
  float c_state[MAXR] = {0.f, 0.f}, h_state[MAXR] = {0.f, 0.f};

  for (int s = 0; s < Tg; ++s) {
    float4 xp[MAXR];
    bool act[MAXR];
    int t_idx[MAXR];
#pragma unroll
    for (int e = 0; e < MAXR; ++e) {
      const int r = rg + e * nrg;
      act[e] = false;
      xp[e] = make_float4(0.f, 0.f, 0.f, 0.f);
      t_idx[e] = 0;
      if (r < REC_ROWS) {
        const int len = s_len[r];
        act[e] = s < len;
        t_idx[e] = dir ? (len - 1 - s) : s;
        if (act[e]) {
          const size_t off = ((size_t)(row0 + r) * T + t_idx[e]) * NX + (size_t)dir * 4 * U + 4 * unit;
          xp[e] = *reinterpret_cast<const float4*>(xproj + off);
        }
      }
    }
    float4 acc[MAXR];
#pragma unroll
    for (int e = 0; e < MAXR; ++e) acc[e] = xp[e];
    if (s > 0) {
      if (tid == 0) wait_counter(ctr, (unsigned)(p.G * s));
      __syncthreads();
      const float* hsrc = hx + (((size_t)((s - 1) & 1) * ndir + dir) * p.Bpad + row0) * U;
      const int chunks_per_row = U / 4;
      for (int i = tid; i < REC_ROWS * chunks_per_row; i += REC_THREADS) {
        const int r = i / chunks_per_row, ch = i % chunks_per_row;
        cp_async16(s_h + r * U + ch * 4, hsrc + (size_t)r * U + ch * 4);
      }
      cp_async_wait_all();
      __syncthreads();
      const float4* wcol = reinterpret_cast<const float4*>(s_w) + ul;
      for (int k = 0; k < U; ++k) {
        const float4 w = wcol[(size_t)k * upc];
#pragma unroll
        for (int e = 0; e < MAXR; ++e) {
          const int r = rg + e * nrg;
          if (r < REC_ROWS) {
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
    for (int e = 0; e < MAXR; ++e) {
      const int r = rg + e * nrg;
      if (r >= REC_ROWS) continue;
      float cn, hn;
      lstm_gates(acc[e].x, acc[e].y, acc[e].z, acc[e].w, c_state[e], cn, hn);
      if (act[e]) {
        c_state[e] = cn;
        h_state[e] = hn;
      }
      const int b = row0 + r;
      if (b < B) {
        hx[(((size_t)(s & 1) * ndir + dir) * p.Bpad + b) * U + unit] = h_state[e];
        if (act[e]) out[(size_t)b * d.out_batch_stride + (size_t)t_idx[e] * (ndir * U) + dir * U + unit] = hn;
      }
    }
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      red_release_add_u32(ctr, 1u);
    }
  }
#pragma unroll
  for (int e = 0; e < MAXR; ++e) {
    const int r = rg + e * nrg;
    const int b = row0 + r;
    if (r < REC_ROWS && b < B) {
      d.c_final[((size_t)dir * B + b) * U + unit] = c_state[e];
      d.h_final[((size_t)dir * B + b) * U + unit] = h_state[e];
    }
  }
}

static int rec_upc(int dtype, int U) {
  if (dtype == PLAS_BF16) return 32;
  int upc = 32;
  while (upc > 8 && (size_t)U * 4 * upc * 4 > 160 * 1024) upc >>= 1;
  while (upc > 4 && U % upc != 0) upc >>= 1;  // narrow layers: U = 16, 24, ...
  return upc;
}

static size_t rec_smem_bytes(int dtype, int U, int upc) {
  if (dtype == PLAS_BF16) return (size_t)U * 256 + (size_t)REC_ROWS * (U + 8) * 2 + REC_ROWS * 32 * 2;
  return (size_t)U * 4 * upc * 4 + (size_t)REC_ROWS * U * 4;
}

int rec_tc_launch(const plas_rec_desc& d, cudaStream_t stream);

static bool rec_force_legacy() {
  const char* e = getenv("PLAS_REC_IMPL");
  return e && strcmp(e, "l2") == 0;
}

template <int KS, int NG>
static int rec_try_cluster(RecArgs a, cudaStream_t stream, bool must_fit_one_wave, bool* launched) {
  const plas_rec_desc& d = a.d;
  constexpr int U = KS * 16, G = U / 32;
  *launched = false;
  const size_t smem_c = (size_t)NG * 2 * REC_ROWS * (U + 8) * 2 + (size_t)REC_ROWS * 32 * 2;
  auto fn = rec_bf16_cluster_kernel<KS, NG>;
  PLAS_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
  if (G > 8) PLAS_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  a.n_groups = (d.B + REC_ROWS - 1) / REC_ROWS;
  a.G = G;
  const int clusters = d.ndir * ((a.n_groups + NG - 1) / NG);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(clusters * G));
  cfg.blockDim = dim3(REC_THREADS);
  cfg.dynamicSmemBytes = smem_c;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)G;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int max_clusters = 0;
  cudaError_t qe = cudaOccupancyMaxActiveClusters(&max_clusters, fn, &cfg);
  if (getenv("PLAS_DEBUG"))
    fprintf(stderr, "[plas] rec cluster path: U=%d NG=%d G=%d clusters=%d max_active_clusters=%d (query %s) smem=%zu\n", U,
            NG, G, clusters, max_clusters, cudaGetErrorString(qe), smem_c);
  if (qe != cudaSuccess || max_clusters < 1) {
    (void)cudaGetLastError();
    return PLAS_OK;
  }
  if (must_fit_one_wave && clusters > max_clusters) return PLAS_OK;
  a.group_offset = 0;
  a.groups_here = a.n_groups;
  PLAS_CUDA(cudaLaunchKernelEx(&cfg, fn, a));
  *launched = true;
  return PLAS_OK;
}

// One group per cluster when all clusters are co-resident (lowest latency); otherwise two
// interleaved groups per cluster (half as many clusters, exchange latency hidden).
// Returns PLAS_OK after a launch, 1 when no cluster shape can be scheduled, <0 on error.
template <int KS>
static int rec_launch_cluster_ks(const RecArgs& a, cudaStream_t stream) {
  bool launched = false;
  const char* force = getenv("PLAS_REC_NG");
  const int fng = force ? atoi(force) : 0;
  int rc;
  if (fng == 0 || fng == 1) {
    rc = rec_try_cluster<KS, 1>(a, stream, fng == 0, &launched);
    if (rc || launched) return rc;
  }
  rc = rec_try_cluster<KS, 2>(a, stream, false, &launched);
  if (rc || launched) return rc;
  return 1;
}

static int rec_launch_cluster(const RecArgs& a, cudaStream_t stream) {
  switch (a.d.U) {
    case 64: return rec_launch_cluster_ks<4>(a, stream);
    case 128: return rec_launch_cluster_ks<8>(a, stream);
    case 256: return rec_launch_cluster_ks<16>(a, stream);
    case 512: return rec_launch_cluster_ks<32>(a, stream);
    default: return 1;  // other widths use the L2-exchange kernel
  }
}

static void rec_ws_layout(const plas_rec_desc& d, size_t* o_ctr, size_t* o_hx, size_t* total) {
  const int n_groups = (d.B + REC_ROWS - 1) / REC_ROWS;
  const size_t esz = d.dtype == PLAS_BF16 ? 2 : 4;
  size_t off = 0;
  *o_ctr = off;
  off += ((size_t)d.ndir * n_groups * 4 + 255) & ~size_t(255);
  *o_hx = off;
  off += ((size_t)2 * d.ndir * n_groups * REC_ROWS * d.U * esz + 255) & ~size_t(255);
  *total = off;
}

}  // namespace plas

using namespace plas;

extern "C" int32_t plas_rec_units_per_cta(int32_t dtype, int32_t U) { return rec_upc(dtype, U); }

extern "C" size_t plas_rec_workspace_bytes(const plas_rec_desc* d) {
  size_t a, b, total;
  rec_ws_layout(*d, &a, &b, &total);
  return total;
}

extern "C" int plas_bilstm_rec_fwd(const plas_rec_desc* d, void* workspace, size_t workspace_bytes,
                                   plas_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PLAS_REQUIRE(d && workspace, "rec: null argument");
  PLAS_REQUIRE(d->dtype == PLAS_F32 || d->dtype == PLAS_BF16, "rec: dtype %d", d->dtype);
  PLAS_REQUIRE(d->B > 0 && d->T > 0 && d->U > 0 && (d->ndir == 1 || d->ndir == 2), "rec: bad shape");
  PLAS_REQUIRE(d->xproj && d->whh && d->lengths && d->out && d->c_final && d->h_final, "rec: null tensor");
  const int upc = rec_upc(d->dtype, d->U);
  if (d->dtype == PLAS_BF16)
    PLAS_REQUIRE(d->U % 64 == 0 && d->U <= 768, "rec(bf16): U=%d must be a multiple of 64 and <= 768", d->U);
  else
    PLAS_REQUIRE(d->U % upc == 0 && d->U % 4 == 0 && d->U <= 1280, "rec(f32): U=%d unsupported (upc=%d)", d->U, upc);
  PLAS_REQUIRE(d->out_batch_stride >= (int64_t)d->T * d->ndir * d->U, "rec: out_batch_stride too small");
  size_t o_ctr, o_hx, total;
  rec_ws_layout(*d, &o_ctr, &o_hx, &total);
  PLAS_REQUIRE(workspace_bytes >= total, "rec: workspace %zu < %zu", workspace_bytes, total);

  RecArgs a;
  a.d = *d;
  a.counters = (unsigned*)((unsigned char*)workspace + o_ctr);
  a.hx = (unsigned char*)workspace + o_hx;
  a.n_groups = (d->B + REC_ROWS - 1) / REC_ROWS;
  a.upc = upc;
  a.G = d->U / upc;
  a.Bpad = a.n_groups * REC_ROWS;

  // fastest path: tcgen05 with W_hh resident in tensor memory (rec_tc.cu); writes the zero padding itself
  {
    const int rc = rec_tc_launch(*d, stream);
    if (rc != 1) return rc;  // 1 = not eligible / not schedulable
  }
  if (!d->out_zeroed)  // the other kernels only write active positions
    PLAS_CUDA(cudaMemsetAsync(d->out, 0, (size_t)d->B * d->out_batch_stride * (d->dtype == PLAS_BF16 ? 2 : 4), stream));
  // next: one thread-block cluster per (direction, group), mma.sync with W_hh resident in registers
  if (d->dtype == PLAS_BF16 && a.G <= 16 && !rec_force_legacy()) {
    int rc = rec_launch_cluster(a, stream);
    if (rc != 1) return rc;  // 1 = cluster shape not schedulable on this device: use the L2-exchange kernel
  }

  PLAS_CUDA(cudaMemsetAsync(a.counters, 0, (size_t)d->ndir * a.n_groups * 4, stream));
  const size_t smem = rec_smem_bytes(d->dtype, d->U, upc);
  PLAS_REQUIRE(smem <= 227 * 1024, "rec: needs %zu bytes of shared memory", smem);
  const void* fn = d->dtype == PLAS_BF16 ? (const void*)rec_bf16_kernel : (const void*)rec_f32_kernel;
  PLAS_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  PLAS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, REC_THREADS, smem));
  const int resident = per_sm * num_sms();
  const int ctas_per_group = d->ndir * a.G;
  const int gpl = resident / ctas_per_group;
  PLAS_REQUIRE(gpl >= 1, "rec: %d CTAs per group cannot be co-resident (%d slots)", ctas_per_group, resident);
  for (int g0 = 0; g0 < a.n_groups; g0 += gpl) {
    a.group_offset = g0;
    a.groups_here = (a.n_groups - g0 < gpl) ? (a.n_groups - g0) : gpl;
    void* args[] = {&a};
    PLAS_CUDA(cudaLaunchCooperativeKernel(fn, dim3(ctas_per_group * a.groups_here), dim3(REC_THREADS), args, smem, stream));
  }
  return PLAS_OK;
}
