// K1 -- fused acoustic front-end for sm_100a.
//
// Replaces calculate_acoustic_features (reference preprocess_all.py:69-130): framing, window,
// mixed-radix FFT power spectrum (n_fft = 400/320 are not powers of two -> real FFT through a
// half-size complex Stockham FFT with radices {2,3,4,5,8}), mel filterbank (sparse rows), log /
// dB, DCT-II, energy column, deltas and the per-channel normalisation of
// utils/dataset_utils.py:213-220.
//
// Work decomposition: one CTA = 32 consecutive frames of one utterance.  The CTA stages the
// contiguous sample span of its frames in shared memory with coalesced loads (frames overlap
// 2.5x, so HBM sees each sample once), then each warp owns a frame at a time: the FFT ping-pongs
// between two per-warp shared buffers and only __syncwarp() separates the passes.
//   speechpy back end : everything is per-frame -> ONE kernel writes the final features.
//   librosa back end  : top_db clips against the utterance-global max, so kernel A writes dB
//                       values + an atomicMax per utterance, kernel B clips (plain MFE: a streaming
//                       float4 kernel, fe_librosa_clip_kernel) or clips + DCTs, kernel C adds the
//                       Savitzky-Golay deltas along time.
// The BASELINE window (n_fft = 400) is a template instance with every size a constant: radix 8-5-5
// passes, window pairs in registers, conflict-free shared-memory traffic.
#include <stdlib.h>

#include "common.cuh"
#include "../../include/plas.h"

namespace plas {

constexpr int FE_FRAMES = 32;
constexpr int FE_WARPS = 8;
constexpr int FE_THREADS = FE_WARPS * 32;

struct FeArgs {
  plas_frontend_desc d;
  const float* wave;
  const int* n_samples;
  long long wave_stride;
  int B;
  float* feats;
  int* n_frames;
  int T_max;
  int C;
  float* db;       // librosa: [B][T_max][n_mels]
  float* rms;      // librosa: [B][T_max]
  unsigned* umax;  // librosa: [B]
  float* base;     // librosa + deltas: [B][T_max][Dbase]
  int generic_fft; // debugging / A-B: run the generic Stockham passes even for n_fft = 400
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)
__device__ __forceinline__ float2 mul_pi(float2 a) { return make_float2(-a.y, a.x); }  // a * (+i)

template <int R> struct Bfly;
template <> struct Bfly<2> {
  static __device__ __forceinline__ void run(float2* v) {
    float2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
  }
};
template <> struct Bfly<3> {
  static __device__ __forceinline__ void run(float2* v) {
    const float s = 0.86602540378443864676f;
    float2 t1 = cadd(v[1], v[2]);
    float2 t2 = make_float2(v[0].x - 0.5f * t1.x, v[0].y - 0.5f * t1.y);
    float2 d = csub(v[1], v[2]);
    float2 t3 = make_float2(s * d.x, s * d.y);
    v[0] = cadd(v[0], t1);
    v[1] = cadd(t2, mul_mi(t3));
    v[2] = cadd(t2, mul_pi(t3));
  }
};
template <> struct Bfly<4> {
  static __device__ __forceinline__ void run(float2* v) {
    float2 a = cadd(v[0], v[2]), b = csub(v[0], v[2]);
    float2 c = cadd(v[1], v[3]), d = csub(v[1], v[3]);
    v[0] = cadd(a, c);
    v[2] = csub(a, c);
    v[1] = cadd(b, mul_mi(d));
    v[3] = cadd(b, mul_pi(d));
  }
};
template <> struct Bfly<5> {
  static __device__ __forceinline__ void run(float2* v) {
    const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
    const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
    float2 a1 = cadd(v[1], v[4]), a2 = cadd(v[2], v[3]);
    float2 b1 = csub(v[1], v[4]), b2 = csub(v[2], v[3]);
    float2 p1 = make_float2(v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y);
    float2 p2 = make_float2(v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y);
    float2 q1 = make_float2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y);
    float2 q2 = make_float2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y);
    v[0] = make_float2(v[0].x + a1.x + a2.x, v[0].y + a1.y + a2.y);
    v[1] = cadd(p1, mul_mi(q1));
    v[4] = cadd(p1, mul_pi(q1));
    v[2] = cadd(p2, mul_mi(q2));
    v[3] = cadd(p2, mul_pi(q2));
  }
};
template <> struct Bfly<8> {
  static __device__ __forceinline__ void run(float2* v) {
    const float h = 0.70710678118654752440f;
    float2 e[4] = {v[0], v[2], v[4], v[6]};
    float2 o[4] = {v[1], v[3], v[5], v[7]};
    Bfly<4>::run(e);
    Bfly<4>::run(o);
    o[1] = make_float2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));   // * (h - ih)
    o[2] = mul_mi(o[2]);
    o[3] = make_float2(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y));  // * (-h - ih)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[k] = cadd(e[k], o[k]);
      v[k + 4] = csub(e[k], o[k]);
    }
  }
};

// One Stockham autosort pass of radix R over n complex points held in shared memory by a warp.
template <int R>
__device__ __forceinline__ void fft_pass(const float2* __restrict__ in, float2* __restrict__ out,
                                         const float2* __restrict__ tw, int n, int Ns, int lane) {
  const int M = n / R;
  const int twstride = n / (Ns * R);
  for (int j = lane; j < M; j += 32) {
    float2 v[R];
    const int k = j % Ns;
#pragma unroll
    for (int t = 0; t < R; ++t) {
      v[t] = in[j + t * M];
      if (t > 0 && Ns > 1) v[t] = cmul(v[t], tw[k * t * twstride]);
    }
    Bfly<R>::run(v);
    const int j0 = (j - k) * R + k;
#pragma unroll
    for (int u = 0; u < R; ++u) out[j0 + u * Ns] = v[u];
  }
}

// n_fft = 400 (window 25 ms, the BASELINE configurations): the 200-point complex FFT as radix 8, 5, 5 Stockham passes with
// every stride, modulus and twiddle step a compile-time constant (the generic passes above spend most of their instructions on
// runtime index arithmetic), 16-byte stores out of the radix-8 pass, and the window applied while the first pass loads its
// samples straight from the staged span.  Result in bufA.
// Returns this lane's share of sum(x^2) over the raw samples when want_sq (librosa's rms energy column), else 0.
// wreg: this lane's eight window pairs win[2 (lane + 25 t)], win[2 (lane + 25 t) + 1], loaded once per CTA (the same for every frame).
__device__ __forceinline__ float fft200_windowed(const float* __restrict__ x, const float2 (&wreg)[8],
                                                 const float2* __restrict__ tw, float2* __restrict__ bufA,
                                                 float2* __restrict__ bufB, int lane, bool want_sq) {
  float sq = 0.f;
  if (lane < 25) {  // pass 1: radix 8, Ns = 1: butterfly j reads elements j + 25 t, writes 8 j .. 8 j + 7
    const float2* x2 = reinterpret_cast<const float2*>(x);
    float2 v[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const float2 xv = x2[lane + 25 * t], wv = wreg[t];
      v[t] = make_float2(xv.x * wv.x, xv.y * wv.y);
      if (want_sq) sq = fmaf(xv.x, xv.x, fmaf(xv.y, xv.y, sq));
    }
    Bfly<8>::run(v);
    // A lane's 8 outputs are one 64-byte block, so the four 16-byte chunks of neighbouring lanes sit 64 bytes apart: stored in
    // the same order by every lane they hit only two of the eight 16-byte bank groups (13 wavefronts per store instead of 4).
    // Lane j therefore starts with chunk (j >> 1) & 3: a register rotation, then instruction i stores chunk (i + rot) & 3.
    float4 c[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) c[u] = make_float4(v[2 * u].x, v[2 * u].y, v[2 * u + 1].x, v[2 * u + 1].y);
    const int rot = (lane >> 1) & 3;
    if (rot & 1) { const float4 t0 = c[0]; c[0] = c[1]; c[1] = c[2]; c[2] = c[3]; c[3] = t0; }
    if (rot & 2) { const float4 t0 = c[0], t1 = c[1]; c[0] = c[2]; c[1] = c[3]; c[2] = t0; c[3] = t1; }
    float4* o = reinterpret_cast<float4*>(bufA + 8 * lane);
#pragma unroll
    for (int u = 0; u < 4; ++u) o[(u + rot) & 3] = c[u];
  }
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 2; ++r) {  // pass 2: radix 5, Ns = 8, twiddle step 5: 40 butterflies
    const int j = lane + 32 * r;
    if (j < 40) {
      const int k = j & 7;
      float2 v[5];
      v[0] = bufA[j];
#pragma unroll
      for (int t = 1; t < 5; ++t) v[t] = cmul(bufA[j + 40 * t], tw[160 + 8 * (t - 1) + k]);  // w200^(5 k t)
      Bfly<5>::run(v);
      float2* o = bufB + (j - k) * 5 + k;
#pragma unroll
      for (int u = 0; u < 5; ++u) o[8 * u] = v[u];
    }
  }
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 2; ++r) {  // pass 3: radix 5, Ns = 40, twiddle step 1
    const int j = lane + 32 * r;
    if (j < 40) {
      float2 v[5];
      v[0] = bufB[j];
#pragma unroll
      for (int t = 1; t < 5; ++t) v[t] = cmul(bufB[j + 40 * t], tw[40 * (t - 1) + j]);  // w200^(j t): contiguous in j, no bank conflicts
      Bfly<5>::run(v);
#pragma unroll
      for (int u = 0; u < 5; ++u) bufA[j + 40 * u] = v[u];
    }
  }
  __syncwarp();
  return sq;
}

__device__ __forceinline__ int frames_of(const plas_frontend_desc& d, int N) {
  if (d.backend == 1) return 1 + N / d.hop;
  return N >= d.n_fft ? (N - d.n_fft) / d.hop : 0;
}

__device__ __forceinline__ float norm_ch(const FeArgs& p, float v, int c) {
  if (p.d.mean) v = (v - p.d.mean[c]) / p.d.stdv[c];
  return v;
}

// 4 CTAs per SM (<= 64 registers, ~55 KB of shared memory each at n_fft = 400): the kernel is issue-bound, occupancy hides its
// shared-memory latencies
// NFFT = 400: the BASELINE window with every size a compile-time constant (fft200_windowed, unrolled unpack, constant buffer
// offsets); NFFT = 0: any even n_fft at run time through the generic Stockham passes
template <int NFFT>
__global__ void __launch_bounds__(FE_THREADS, 4) fe_spectral_kernel(FeArgs p) {
  extern __shared__ __align__(16) unsigned char fe_smem[];
  const plas_frontend_desc& d = p.d;
  const int n_fft = NFFT ? NFFT : d.n_fft, hop = d.hop, n = n_fft / 2, n_mels = d.n_mels;
  const int b = blockIdx.y;
  const int tile0 = blockIdx.x * FE_FRAMES;
  const int N = p.n_samples[b];
  const bool librosa = d.backend == 1;
  const int T_b = min(frames_of(d, N), p.T_max);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (blockIdx.x == 0 && tid == 0) p.n_frames[b] = T_b;

  if (tile0 >= T_b) {
    if (!librosa) {  // zero-fill the padded tail (tf padded_batch pads with 0.0)
      const int rows = min(FE_FRAMES, p.T_max - tile0);
      float* dst = p.feats + ((size_t)b * p.T_max + tile0) * p.C;
      for (int i = tid; i < rows * p.C; i += FE_THREADS) dst[i] = 0.f;
    }
    return;
  }

  // ---- shared memory carve-up --------------------------------------------------------
  // byte offsets from the one shared base, so that every pointer below keeps the shared address space (LDS / STS, not generic
  // loads); the filterbank weights and the span are 16-byte aligned (float4 rows / float4 staging), see fe_spectral_smem
  const int span_len = (FE_FRAMES - 1) * hop + n_fft;
  const int off_twu = 8 * n;
  const int off_win = off_twu + 8 * (n + 1);
  const int off_fbw = (off_win + 4 * n_fft + 15) & ~15;
  const int off_fbs = off_fbw + 4 * d.fb_total;
  const int off_span = (off_fbs + 12 * n_mels + 15) & ~15;
  const int off_work = (off_span + 4 * span_len + 15) & ~15;
  const int work_stride = 2 * n + 2;  // float2 elements per warp (two ping-pong buffers)
  float2* s_tw = reinterpret_cast<float2*>(fe_smem);
  float2* s_twu = reinterpret_cast<float2*>(fe_smem + off_twu);
  float* s_win = reinterpret_cast<float*>(fe_smem + off_win);
  float* s_fbw = reinterpret_cast<float*>(fe_smem + off_fbw);
  int* s_fbs = reinterpret_cast<int*>(fe_smem + off_fbs);
  float* s_span = reinterpret_cast<float*>(fe_smem + off_span);
  float2* s_work = reinterpret_cast<float2*>(fe_smem + off_work) + warp * work_stride;

  if constexpr (NFFT == 400) {
    // the twiddles of the two radix-5 passes in the order their lanes read them: [0, 160) w^(j t) as [t - 1][j], j < 40 (pass 3);
    // [160, 192) w^(5 k t) as [t - 1][k], k < 8 (pass 2).  (Read as tw[j t], strides 2 and 4 replayed a quarter of pass 3's wavefronts.)
    for (int i = tid; i < 192; i += FE_THREADS) {
      const int idx = i < 160 ? (i % 40) * (i / 40 + 1) : 5 * ((i - 160) & 7) * (((i - 160) >> 3) + 1);
      s_tw[i] = reinterpret_cast<const float2*>(d.tw)[idx];
    }
  } else {
    for (int i = tid; i < n; i += FE_THREADS) s_tw[i] = reinterpret_cast<const float2*>(d.tw)[i];
  }
  for (int i = tid; i <= n; i += FE_THREADS) s_twu[i] = reinterpret_cast<const float2*>(d.tw_unpack)[i];
  for (int i = tid; i < n_fft; i += FE_THREADS) s_win[i] = d.window[i];
  for (int i = tid; i < d.fb_total; i += FE_THREADS) s_fbw[i] = d.fb_w[i];
  for (int i = tid; i < n_mels; i += FE_THREADS) {
    s_fbs[i] = d.fb_start[i];
    s_fbs[n_mels + i] = d.fb_len[i];
    s_fbs[2 * n_mels + i] = d.fb_off[i];
  }
  {
    const float* w = p.wave + (size_t)b * p.wave_stride;
    const int base = tile0 * hop;
    const int first = librosa ? base - n : base;
    if (first >= 0 && first + span_len <= N && (span_len & 3) == 0 && (reinterpret_cast<uintptr_t>(w + first) & 15) == 0) {
      // interior tile (no reflection, no tail): 16-byte loads
      const float4* w4 = reinterpret_cast<const float4*>(w + first);
      float4* s4 = reinterpret_cast<float4*>(s_span);
      for (int i = tid; i < (span_len >> 2); i += FE_THREADS) s4[i] = __ldg(w4 + i);
    } else
    for (int i = tid; i < span_len; i += FE_THREADS) {
      int src = base + i;
      if (librosa) {  // centre=True, pad_mode='reflect' by n_fft/2
        src -= n;
        if (src < 0) src = -src;
        if (src >= N) src = 2 * (N - 1) - src;
        src = max(0, min(src, N - 1));
      }
      s_span[i] = (src < N) ? w[src] : 0.f;
    }
  }
  __syncthreads();

  const float eps64 = 2.220446049250313e-16f;  // np.finfo(float).eps (speechpy zero_handling)
  float2 wreg[8];  // n_fft = 400: the window pairs of this lane's radix-8 butterfly stay in registers across the CTA's frames
  if constexpr (NFFT == 400) {
#pragma unroll
    for (int t = 0; t < 8; ++t) wreg[t] = lane < 25 ? reinterpret_cast<const float2*>(s_win)[lane + 25 * t] : make_float2(0.f, 0.f);
  }
  for (int f = warp; f < FE_FRAMES; f += FE_WARPS) {
    const int t = tile0 + f;
    if (t >= p.T_max) break;
    if (t >= T_b) {
      if (!librosa) {
        float* dst = p.feats + ((size_t)b * p.T_max + t) * p.C;
        for (int i = lane; i < p.C; i += 32) dst[i] = 0.f;
      }
      continue;
    }
    const float* x = s_span + f * hop;
    float2* bufA = s_work;
    float2* bufB = s_work + n + 1;

    float rms = 0.f;
    const bool want_rms = librosa && d.energy;
    if (NFFT != 400 && want_rms) {  // (the n_fft = 400 path sums the squares while its first FFT pass loads the samples)
      float s = 0.f;
      for (int i = lane; i < n_fft; i += 32) s += x[i] * x[i];
      rms = sqrtf(warp_sum(s) / (float)n_fft);
    }
    float2* src = bufA;
    float2* dst = bufB;
    if constexpr (NFFT == 400) {
      const float sq = fft200_windowed(x, wreg, s_tw, bufA, bufB, lane, want_rms);
      if (want_rms) rms = sqrtf(warp_sum(sq) / (float)n_fft);
    } else {
      for (int j = lane; j < n; j += 32)
        bufA[j] = make_float2(x[2 * j] * s_win[2 * j], x[2 * j + 1] * s_win[2 * j + 1]);
      __syncwarp();
    }
    int Ns = 1;
    for (int s = 0; s < (NFFT == 400 ? 0 : d.n_fac); ++s) {
      const int R = d.fac[s];
      switch (R) {
        case 2: fft_pass<2>(src, dst, s_tw, n, Ns, lane); break;
        case 3: fft_pass<3>(src, dst, s_tw, n, Ns, lane); break;
        case 4: fft_pass<4>(src, dst, s_tw, n, Ns, lane); break;
        case 5: fft_pass<5>(src, dst, s_tw, n, Ns, lane); break;
        default: fft_pass<8>(src, dst, s_tw, n, Ns, lane); break;
      }
      __syncwarp();
      float2* tmp = src; src = dst; dst = tmp;
      Ns *= R;
    }
    // src = Z (half-size complex spectrum).  Unpack to the real spectrum's power, bins 0..n.
    float* P = reinterpret_cast<float*>(dst);
    // the halves of xe / xo are folded into the scale of the power: |X|^2 = |2 xe + w^k 2 xo|^2 / 4
    const float pscale = librosa ? 0.25f : 0.25f / (float)n_fft;
    float esum = 0.f;
    // bins k and n - k come from the same two points of Z: X[k] = xe + w^k xo, X[n-k] = conj(xe - w^k xo)
#pragma unroll
    for (int k = lane; 2 * k <= n; k += 32) {
      const float2 zk = src[k];
      const float2 zm = src[k == 0 ? 0 : n - k];
      const float2 xe = make_float2(zk.x + zm.x, zk.y - zm.y);
      const float2 xo = make_float2(zk.y + zm.y, zm.x - zk.x);
      const float2 tx = cmul(s_twu[k], xo);
      const float2 Xa = cadd(xe, tx), Xb = csub(xe, tx);
      const float pa = (Xa.x * Xa.x + Xa.y * Xa.y) * pscale;
      const float pb = (Xb.x * Xb.x + Xb.y * Xb.y) * pscale;
      P[k] = pa;
      P[n - k] = pb;
      esum += (2 * k == n) ? pa : pa + pb;
    }
    if (lane < 3) P[n + 1 + lane] = 0.f;  // read (times a zero weight) by the four-wide filterbank rows
    __syncwarp();
    const float E = warp_sum(esum);

    float* mel = reinterpret_cast<float*>(src);
    // librosa: the dB value goes straight from the filterbank accumulator to HBM (top_db is applied in fe_librosa_post_kernel)
    float* db = librosa ? p.db + ((size_t)b * p.T_max + t) * n_mels : nullptr;
    float mx = -INFINITY;
    for (int m = lane; m < n_mels; m += 32) {
      const int st = s_fbs[m], len = s_fbs[n_mels + m], o = s_fbs[2 * n_mels + m];
      float acc = 0.f;
      if (((len | o) & 3) == 0) {  // rows zero-padded to multiples of four weights by the host (plas.h)
        const float4* w4 = reinterpret_cast<const float4*>(s_fbw + o);
        const float* Pm = P + st;
#pragma unroll 1
        for (int i = 0; i < len; i += 4) {  // 1 to 4 trips: an unrolled loop's prologue / remainder code costs more than it saves
          const float4 w = w4[i >> 2];
          acc = fmaf(w.x, Pm[i], acc);
          acc = fmaf(w.y, Pm[i + 1], acc);
          acc = fmaf(w.z, Pm[i + 2], acc);
          acc = fmaf(w.w, Pm[i + 3], acc);
        }
      } else {
        for (int i = 0; i < len; ++i) acc = fmaf(s_fbw[o + i], P[st + i], acc);
      }
      if (librosa) {
        const float v = d.feature_type == 0 ? acc * acc : acc;  // amplitude_to_db squares its input (preprocess_all.py:83)
        const float val = 3.01029995663981195f * __log2f(fmaxf(1e-10f, v));  // 10 log10: lg2.approx (rel. error 2^-22) * 10 log10(2)
        db[m] = val;
        mx = fmaxf(mx, val);
      } else {
        mel[m] = acc;
      }
    }
    __syncwarp();

    if (!librosa) {
      float* out = p.feats + ((size_t)b * p.T_max + t) * p.C;
      // speechpy extract_derivative_feature (preprocess_all.py:120-123): two passes of derivative_extraction along the
      // FEATURE axis over the D base features in cep[0..D), then [f_k, d f_k, dd f_k] interleaved per base channel
      auto sp_deltas_out = [&](float* cep, int D) {
        for (int pass = 0; pass < 2; ++pass) {
          const float* F = cep + pass * D;
          float* G = cep + (pass + 1) * D;
          for (int k = lane; k < D; k += 32) {
            float acc = F[min(k + 1, D - 1)] + 2.f * F[min(k + 2, D - 1)];
            if (!d.sp_delta_literal) acc -= F[max(k - 1, 0)] + F[max(k - 2, 0)];
            G[k] = acc / 10.f;
          }
          __syncwarp();
        }
        for (int i = lane; i < 3 * D; i += 32) {
          const int k = i / 3, which = i - 3 * k;
          out[i] = norm_ch(p, cep[which * D + k], i);
        }
      };
      if (d.feature_type == 0) {  // speechpy mfe: log([mel, energy] + 1e-8)
        float* cep = P;  // [3][n_mels + 1] with --deltas (the power spectrum is dead by now)
        for (int m = lane; m < n_mels; m += 32) {
          float v = mel[m];
          v = (v == 0.f) ? eps64 : v;
          v = logf(v + 1e-8f);
          if (d.deltas) cep[m] = v;
          else out[m] = norm_ch(p, v, m);
        }
        if (lane == 0) {
          float v = (E == 0.f) ? eps64 : E;
          v = logf(v + 1e-8f);
          if (d.deltas) cep[n_mels] = v;
          else out[n_mels] = norm_ch(p, v, n_mels);
        }
        if (d.deltas) {
          __syncwarp();
          sp_deltas_out(cep, n_mels + 1);
        }
      } else {  // speechpy mfcc: DCT-II(log mel), c0 := log(energy), optional feature-axis deltas
        for (int m = lane; m < n_mels; m += 32) {
          const float v = mel[m];
          mel[m] = logf((v == 0.f) ? eps64 : v);
        }
        __syncwarp();
        const int D = d.n_mfcc;
        float* cep = P;  // [3][D]
        for (int q = lane; q < D; q += 32) {
          float acc = 0.f;
          const float* row = d.dct + (size_t)q * n_mels;
          for (int m = 0; m < n_mels; ++m) acc = fmaf(__ldg(row + m), mel[m], acc);
          if (q == 0) acc = logf((E == 0.f) ? eps64 : E);
          cep[q] = acc;
        }
        __syncwarp();
        if (d.deltas) {
          sp_deltas_out(cep, D);
        } else {
          for (int k = lane; k < D; k += 32) out[k] = norm_ch(p, cep[k], k);
        }
      }
    } else {
      // librosa: utterance max of the dB values written above
      mx = warp_max(mx);
      if (lane == 0) {
        atomicMax(p.umax + b, float_to_ordered(mx));
        if (d.energy) p.rms[(size_t)b * p.T_max + t] = rms;
      }
    }
    __syncwarp();
  }
}

// librosa kernel B for plain MFE (no deltas, no energy column, n_mels a multiple of 4 -- the c2 / c5 mfe80 case): the top_db
// clip and the normalisation are elementwise over [T_max][n_mels], so this is a streaming kernel of 16-byte accesses with four
// independent loads in flight per thread (the warp-per-frame kernel below moved 1.7 TB/s on it).
constexpr int FE_CLIP_VEC = 4;  // float4 per thread
__global__ void __launch_bounds__(256) fe_librosa_clip_kernel(FeArgs p) {
  const plas_frontend_desc& d = p.d;
  const int b = blockIdx.y;
  const int q4 = d.n_mels >> 2;                  // float4 per frame
  const int total = p.T_max * q4;                // float4 per utterance
  const int valid = min(frames_of(d, p.n_samples[b]), p.T_max) * q4;
  const float floor_db = ordered_to_float(p.umax[b]) - 80.0f;
  const float4* in = reinterpret_cast<const float4*>(p.db + (size_t)b * p.T_max * d.n_mels);
  float4* out = reinterpret_cast<float4*>(p.feats + (size_t)b * p.T_max * d.n_mels);
  const int e0 = blockIdx.x * (256 * FE_CLIP_VEC) + threadIdx.x;
  float4 v[FE_CLIP_VEC];
#pragma unroll
  for (int u = 0; u < FE_CLIP_VEC; ++u) {
    const int e = e0 + u * 256;
    v[u] = e < valid ? __ldcs(in + e) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int u = 0; u < FE_CLIP_VEC; ++u) {
    const int e = e0 + u * 256;
    if (e >= total) break;
    float4 r = v[u];
    if (e < valid) {
      r = make_float4(fmaxf(r.x, floor_db), fmaxf(r.y, floor_db), fmaxf(r.z, floor_db), fmaxf(r.w, floor_db));
      if (d.mean) {
        const int c = (e % q4) * 4;
        const float4 mu = *reinterpret_cast<const float4*>(d.mean + c), sd = *reinterpret_cast<const float4*>(d.stdv + c);
        r = make_float4((r.x - mu.x) / sd.x, (r.y - mu.y) / sd.y, (r.z - mu.z) / sd.z, (r.w - mu.w) / sd.w);
      }
    }
    out[e] = r;  // frames past the utterance's end are zero (tf padded_batch)
  }
}

// librosa kernel B: top_db clip (+ DCT for mfcc, + rms column); writes final features when
// there are no deltas, otherwise the base features for kernel C.
__global__ void __launch_bounds__(FE_THREADS, 4) fe_librosa_post_kernel(FeArgs p) {
  extern __shared__ __align__(16) unsigned char fe_smem[];
  const plas_frontend_desc& d = p.d;
  const int n_mels = d.n_mels;
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int T_b = min(frames_of(d, p.n_samples[b]), p.T_max);
  const int nb = (d.feature_type == 0 ? n_mels : d.n_mfcc);
  const int Dbase = nb + (d.energy ? 1 : 0);
  const bool use_delta = d.deltas && T_b >= 9;
  float* s_db = reinterpret_cast<float*>(fe_smem) + warp * n_mels;
  float* s_dct = reinterpret_cast<float*>(fe_smem) + FE_WARPS * n_mels;  // [n_mfcc][n_mels | 1] (odd stride: no bank conflicts across q)
  const int dct_stride = n_mels | 1;
  const bool dct_split = 2 * nb <= 32;
  if (d.feature_type != 0) {
    for (int i = threadIdx.x; i < nb * n_mels; i += FE_THREADS) s_dct[(i / n_mels) * dct_stride + i % n_mels] = d.dct[i];
    __syncthreads();
  }
  const float floor_db = ordered_to_float(p.umax[b]) - 80.0f;
  for (int f = warp; f < FE_FRAMES; f += FE_WARPS) {
    const int t = blockIdx.x * FE_FRAMES + f;
    if (t >= p.T_max) break;
    float* out = p.feats + ((size_t)b * p.T_max + t) * p.C;
    if (t >= T_b) {
      for (int i = lane; i < p.C; i += 32) out[i] = 0.f;
      continue;
    }
    const float* db = p.db + ((size_t)b * p.T_max + t) * n_mels;
    float* base = use_delta ? p.base + ((size_t)b * p.T_max + t) * Dbase : nullptr;
    const float rms = d.energy ? p.rms[(size_t)b * p.T_max + t] : 0.f;
    if (d.feature_type == 0) {
      for (int m = lane; m < n_mels; m += 32) {
        const float v = fmaxf(db[m], floor_db);
        if (use_delta) base[m] = v;
        else if (d.deltas) { out[3 * m] = norm_ch(p, v, 3 * m); out[3 * m + 1] = norm_ch(p, 0.f, 3 * m + 1); out[3 * m + 2] = norm_ch(p, 0.f, 3 * m + 2); }
        else out[m] = norm_ch(p, v, m);
      }
    } else {
      for (int m = lane; m < n_mels; m += 32) s_db[m] = fmaxf(db[m], floor_db);
      __syncwarp();
      if (dct_split) {
        // 2 nb <= 32: lanes [0, nb) take the first half of the mel axis of coefficient q = lane, lanes [nb, 2 nb) the second
        // half (13 of 32 lanes walking all 40 mels was 2/3 of this kernel's instructions); DCT rows from shared memory
        const int part = lane >= nb, q = lane - part * nb, mh = (n_mels + 1) >> 1;
        float acc = 0.f;
        if (lane < 2 * nb) {
          const float* row = s_dct + q * dct_stride;
          const int m1 = part ? n_mels : mh;
#pragma unroll 2
          for (int m = part ? mh : 0; m < m1; ++m) acc = fmaf(row[m], s_db[m], acc);
        }
        acc += __shfl_down_sync(0xffffffffu, acc, nb);
        if (lane < nb) {
          if (use_delta) base[q] = acc;
          else if (d.deltas) { out[3 * q] = norm_ch(p, acc, 3 * q); out[3 * q + 1] = norm_ch(p, 0.f, 3 * q + 1); out[3 * q + 2] = norm_ch(p, 0.f, 3 * q + 2); }
          else out[q] = norm_ch(p, acc, q);
        }
      } else
      for (int q = lane; q < nb; q += 32) {
        float acc = 0.f;
        const float* row = s_dct + q * dct_stride;
        for (int m = 0; m < n_mels; ++m) acc = fmaf(row[m], s_db[m], acc);
        if (use_delta) base[q] = acc;
        else if (d.deltas) { out[3 * q] = norm_ch(p, acc, 3 * q); out[3 * q + 1] = norm_ch(p, 0.f, 3 * q + 1); out[3 * q + 2] = norm_ch(p, 0.f, 3 * q + 2); }
        else out[q] = norm_ch(p, acc, q);
      }
      __syncwarp();
    }
    if (d.energy && lane == 0) {
      if (use_delta) base[nb] = rms;
      else if (d.deltas) { out[3 * nb] = norm_ch(p, rms, 3 * nb); out[3 * nb + 1] = norm_ch(p, 0.f, 3 * nb + 1); out[3 * nb + 2] = norm_ch(p, 0.f, 3 * nb + 2); }
      else out[nb] = norm_ch(p, rms, nb);
    }
  }
}

// librosa kernel C: Savitzky-Golay deltas along time (width 9; order-1 and order-2 taps; the
// first/last four frames take the value at the nearest interior centre = scipy mode='interp'),
// interleaved [f, d, dd] per base channel (preprocess_all.py:125-129).
__global__ void fe_librosa_delta_kernel(FeArgs p) {
  const plas_frontend_desc& d = p.d;
  const int nb = (d.feature_type == 0 ? d.n_mels : d.n_mfcc);
  const int Dbase = nb + (d.energy ? 1 : 0);
  const size_t total = (size_t)p.B * p.T_max * Dbase;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % Dbase);
    const int t = (int)((i / Dbase) % p.T_max);
    const int b = (int)(i / ((size_t)Dbase * p.T_max));
    const int T_b = min(frames_of(d, p.n_samples[b]), p.T_max);
    if (t >= T_b || T_b < 9) continue;  // tail zero-filled / no-delta case handled by kernel B
    const int tc = max(4, min(t, T_b - 5));
    const float* col = p.base + (size_t)b * p.T_max * Dbase + k;
    const float w2[9] = {28.f, 7.f, -8.f, -17.f, -20.f, -17.f, -8.f, 7.f, 28.f};
    float d1 = 0.f, d2 = 0.f;
#pragma unroll
    for (int j = -4; j <= 4; ++j) {
      const float v = col[(size_t)(tc + j) * Dbase];
      d1 = fmaf((float)j, v, d1);
      d2 = fmaf(w2[j + 4], v, d2);
    }
    float* out = p.feats + ((size_t)b * p.T_max + t) * p.C + 3 * k;
    out[0] = norm_ch(p, col[(size_t)t * Dbase], 3 * k);
    out[1] = norm_ch(p, d1 / 60.f, 3 * k + 1);
    out[2] = norm_ch(p, d2 / 462.f, 3 * k + 2);
  }
}

static size_t fe_spectral_smem(const plas_frontend_desc& d) {  // the carve-up of fe_spectral_kernel
  const int n = d.n_fft / 2;
  size_t off = (size_t)8 * n + (size_t)8 * (n + 1) + (size_t)4 * d.n_fft;
  off = (off + 15) & ~size_t(15);
  off += (size_t)4 * d.fb_total + (size_t)12 * d.n_mels;
  off = (off + 15) & ~size_t(15);
  off += (size_t)4 * ((FE_FRAMES - 1) * d.hop + d.n_fft);
  off = (off + 15) & ~size_t(15);
  return off + (size_t)FE_WARPS * (2 * n + 2) * 8;
}

static void fe_ws_layout(const plas_frontend_desc& d, int B, int T_max, size_t* o_db, size_t* o_rms,
                         size_t* o_umax, size_t* o_base, size_t* total) {
  const size_t rows = (size_t)B * T_max;
  const int nb = (d.feature_type == 0 ? d.n_mels : d.n_mfcc);
  const int Dbase = nb + (d.energy ? 1 : 0);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~size_t(255); return o; };
  *o_umax = take((size_t)B * 4);
  *o_db = take(rows * d.n_mels * 4);
  *o_rms = take(rows * 4);
  *o_base = take(d.deltas ? rows * Dbase * 4 : 0);
  *total = d.backend == 1 ? off : 256;
}

}  // namespace plas

using namespace plas;

extern "C" size_t plas_frontend_workspace_bytes(const plas_frontend_desc* d, int32_t B, int32_t T_max) {
  size_t a, b, c, e, total;
  fe_ws_layout(*d, B, T_max, &a, &b, &c, &e, &total);
  return total;
}

extern "C" int plas_frontend_fwd(const plas_frontend_desc* d, const float* wave, const int32_t* n_samples,
                                 int32_t B, int64_t wave_stride, float* feats, int32_t* n_frames,
                                 int32_t T_max, int32_t C, void* workspace, size_t workspace_bytes,
                                 plas_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PLAS_REQUIRE(d && wave && n_samples && feats && n_frames, "frontend: null argument");
  PLAS_REQUIRE(B > 0 && T_max > 0, "frontend: B=%d T_max=%d", B, T_max);
  PLAS_REQUIRE(d->n_fft % 2 == 0 && d->n_fft >= 16 && d->hop > 0, "frontend: n_fft=%d hop=%d", d->n_fft, d->hop);
  int prod = 1;
  PLAS_REQUIRE(d->n_fac >= 1 && d->n_fac <= 8, "frontend: n_fac=%d", d->n_fac);
  for (int i = 0; i < d->n_fac; ++i) {
    const int r = d->fac[i];
    PLAS_REQUIRE(r == 2 || r == 3 || r == 4 || r == 5 || r == 8, "frontend: unsupported radix %d", r);
    prod *= r;
  }
  PLAS_REQUIRE(prod == d->n_fft / 2, "frontend: radices multiply to %d, need %d", prod, d->n_fft / 2);
  const int nb = (d->feature_type == 0 ? d->n_mels : d->n_mfcc);
  int C_expect;
  if (d->backend == 0) C_expect = ((d->feature_type == 0) ? d->n_mels + 1 : d->n_mfcc) * (d->deltas ? 3 : 1);
  else C_expect = (nb + (d->energy ? 1 : 0)) * (d->deltas ? 3 : 1);
  PLAS_REQUIRE(C == C_expect, "frontend: C=%d but the flags produce %d channels", C, C_expect);
  PLAS_REQUIRE(d->feature_type == 0 || d->dct, "frontend: mfcc needs a DCT matrix");
  PLAS_REQUIRE(3 * (nb + (d->feature_type == 0 ? 1 : 0)) <= d->n_fft + 2 || d->backend == 1,
               "frontend: too many base features for the per-warp work buffer");

  FeArgs a;
  a.d = *d;
  a.wave = wave; a.n_samples = n_samples; a.wave_stride = wave_stride; a.B = B;
  a.feats = feats; a.n_frames = n_frames; a.T_max = T_max; a.C = C;
  a.db = nullptr; a.rms = nullptr; a.umax = nullptr; a.base = nullptr;
  a.generic_fft = getenv("PLAS_FE_GENERIC") ? 1 : 0;
  size_t o_db, o_rms, o_umax, o_base, total;
  fe_ws_layout(*d, B, T_max, &o_db, &o_rms, &o_umax, &o_base, &total);
  if (d->backend == 1) {
    PLAS_REQUIRE(workspace && workspace_bytes >= total, "frontend: workspace %zu < %zu", workspace_bytes, total);
    unsigned char* ws = (unsigned char*)workspace;
    a.umax = (unsigned*)(ws + o_umax);
    a.db = (float*)(ws + o_db);
    a.rms = (float*)(ws + o_rms);
    a.base = (float*)(ws + o_base);
    PLAS_CUDA(cudaMemsetAsync(a.umax, 0, (size_t)B * 4, stream));
  }
  const size_t smem = fe_spectral_smem(*d);
  PLAS_REQUIRE(smem <= 227 * 1024, "frontend: %zu bytes of shared memory needed", smem);
  const bool fast400 = d->n_fft == 400 && (d->hop & 1) == 0 && !a.generic_fft;
  auto* spectral = fast400 ? fe_spectral_kernel<400> : fe_spectral_kernel<0>;
  PLAS_CUDA(cudaFuncSetAttribute(spectral, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // gridDim.y carries the utterance index (<= 65535): large batches (the 64 k-utterance sweep of BASELINE configs[4]) go in chunks
  const int Dbase = nb + (d->energy ? 1 : 0);
  constexpr int FE_MAX_B = 32768;
  const FeArgs a0 = a;
  for (int b0 = 0; b0 < B; b0 += FE_MAX_B) {
    const int nb_here = B - b0 < FE_MAX_B ? B - b0 : FE_MAX_B;
    a = a0;
    a.B = nb_here;
    a.wave = a0.wave + (size_t)b0 * wave_stride;
    a.n_samples = a0.n_samples + b0;
    a.feats = a0.feats + (size_t)b0 * T_max * C;
    a.n_frames = a0.n_frames + b0;
    if (d->backend == 1) {
      a.umax = a0.umax + b0;
      a.db = a0.db + (size_t)b0 * T_max * d->n_mels;
      a.rms = a0.rms + (size_t)b0 * T_max;
      if (a0.base) a.base = a0.base + (size_t)b0 * T_max * Dbase;
    }
    dim3 grid((T_max + FE_FRAMES - 1) / FE_FRAMES, nb_here);
    spectral<<<grid, FE_THREADS, smem, stream>>>(a);
    PLAS_CUDA(cudaGetLastError());
    if (d->backend == 1) {
      if (d->feature_type == 0 && !d->deltas && !d->energy && (d->n_mels & 3) == 0 &&
          (!d->mean || (((uintptr_t)d->mean | (uintptr_t)d->stdv) & 15) == 0) && (((uintptr_t)a.db | (uintptr_t)a.feats) & 15) == 0) {
        const int per_block = 256 * FE_CLIP_VEC;
        dim3 grid_c((T_max * (d->n_mels / 4) + per_block - 1) / per_block, nb_here);
        fe_librosa_clip_kernel<<<grid_c, 256, 0, stream>>>(a);
      } else {
        const size_t smem_b = (size_t)FE_WARPS * d->n_mels * 4 + (d->feature_type != 0 ? (size_t)d->n_mfcc * (d->n_mels | 1) * 4 : 0);
        PLAS_REQUIRE(smem_b <= 48 * 1024, "frontend: %zu bytes of shared memory for the DCT rows", smem_b);
        fe_librosa_post_kernel<<<grid, FE_THREADS, smem_b, stream>>>(a);
      }
      PLAS_CUDA(cudaGetLastError());
      if (d->deltas) {
        const size_t total_el = (size_t)nb_here * T_max * Dbase;
        int blocks = (int)((total_el + 255) / 256 < 148 * 16 ? (total_el + 255) / 256 : 148 * 16);
        fe_librosa_delta_kernel<<<blocks, 256, 0, stream>>>(a);
        PLAS_CUDA(cudaGetLastError());
      }
    }
  }
  return PLAS_OK;
}
