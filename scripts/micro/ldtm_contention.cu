// Does tcgen05.ld wait for tensor-core work in flight?  Warp 1 issues chains of 32 TS-form MMAs (A from tensor memory, M = 128,
// N = 16, K = 16 each -- the recurrence's step product) back to back, or idles; warp 2 times tcgen05.ld.16x128b x2 + wait::ld of an
// accumulator the MMAs do not touch.  Reported: average cycles per load pair, idle vs under MMA load, and the MMA chain time
// with and without the concurrent loads.
#include <cstdio>
#include "../../phones_las_b200/csrc/common.cuh"
#include "../../phones_las_b200/csrc/tcgen05.cuh"
using namespace plas;

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void ld_16x128b(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
}

__global__ void __launch_bounds__(128, 1) k(long long* out, int mma_on, int ld_on, int reps) {
  extern __shared__ unsigned char raw_[];
  const uint32_t raw = smem_u32(raw_);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* smem = raw_ + (base - raw);
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ volatile int s_stop;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 8 * 16 * 128 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (tid == 0) { mbar_init(smem_u32(&s_bar), 1); s_stop = 0; }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = s_tmem;
  if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, 16);
      const uint64_t b0 = umma_smem_desc(base);
      uint32_t par = 0;
      long long t0 = clock64();
      int n = 0;
      while (mma_on && (ld_on ? !s_stop : n < reps)) {
#pragma unroll
        for (int i = 0; i < 32; ++i) umma_ts(tb, tb + 128 + 8 * i, b0 + (uint64_t)((i >> 2) * 128 + 2 * (i & 3)), idesc, i != 0);
        umma_commit(smem_u32(&s_bar));
        mbar_wait(smem_u32(&s_bar), par);
        par ^= 1u;
        tc_fence_after();
        ++n;
      }
      out[2 * blockIdx.x + 1] = n ? (clock64() - t0) / n : 0;
    }
    __syncwarp();
  } else if (warp == 2) {
    long long tot = 0;
    if (ld_on) {
      uint32_t r[4];
      for (int it = 0; it < reps; ++it) {
        const long long t0 = clock64();
        ld_16x128b(tb + (64u << 16) + 64u, r);       // lanes 64..79 / 80..95 of this warp's quadrant, columns 64..67: not an MMA target
        ld_16x128b(tb + (80u << 16) + 64u, r + 2);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        tot += clock64() - t0;
        for (int w = 0; w < 20; ++w) asm volatile("nanosleep.u32 20;");  // let the other warp run
      }
      s_stop = 1;
    }
    if ((tid & 31) == 0) out[2 * blockIdx.x] = ld_on ? tot / reps : 0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

int main() {
  long long* out; cudaMalloc(&out, 148 * 16);
  const int smem = 1024 + 8 * 16 * 128;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int mode = 0; mode < 3; ++mode) {
    const int mma_on = mode != 0, ld_on = mode != 2;
    k<<<148, 128, smem>>>(out, mma_on, ld_on, 2000);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[296]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    double ld = 0, mm = 0; for (int i = 0; i < 148; ++i) { ld += h[2 * i]; mm += h[2 * i + 1]; }
    printf("MMA chains %s, loads %s: tcgen05.ld pair + wait %.0f cycles, chain of 32 TS MMAs + commit + wait %.0f cycles\n", mma_on ? "running" : "off", ld_on ? "on" : "off", ld / 148, mm / 148);
  }
  return 0;
}
