// Micro-benchmark: cost of a software grid barrier across 128 co-resident CTAs (288 threads each), variants.
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
__device__ __forceinline__ unsigned ld_acq(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_rlx(const unsigned* p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void red_rel(unsigned* p, unsigned v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void cl_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }

template <int MODE>
__global__ void k(unsigned* bar, unsigned* flags, int iters, int csize) {
  unsigned epoch = 0;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {  // flat: every CTA red + spin on one counter
      __syncthreads();
      if (threadIdx.x == 0) { epoch++; red_rel(bar, 1u); { unsigned sp = 0; while (ld_acq(bar) < epoch * gridDim.x) { if (++sp > (1u << 24)) __trap(); } } }
      __syncthreads();
    } else if (MODE == 1) {  // hierarchical: cluster barrier, rank 0 does the global part, cluster barrier
      cl_sync();
      if (blockIdx.x % csize == 0 && threadIdx.x == 0) { epoch++; red_rel(bar, 1u); { unsigned sp = 0; while (ld_acq(bar) < epoch * (gridDim.x / csize)) { if (++sp > (1u << 24)) __trap(); } } }
      cl_sync();
    } else if (MODE == 2) {  // flat, relaxed polling + one acquire at the end
      __syncthreads();
      if (threadIdx.x == 0) { epoch++; red_rel(bar, 1u); { unsigned sp = 0; while (ld_rlx(bar) < epoch * gridDim.x) { if (++sp > (1u << 24)) __trap(); } } (void)ld_acq(bar); }
      __syncthreads();
    } else if (MODE == 3) {  // per-CTA flags: CTA 0 collects, then broadcasts a generation word
      __syncthreads();
      if (threadIdx.x == 0) { epoch++; asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + blockIdx.x * 32), "r"(epoch) : "memory"); }
      if (blockIdx.x == 0) {
        if (threadIdx.x < gridDim.x) { { unsigned sp = 0; while (ld_acq(flags + threadIdx.x * 32) < (unsigned)(it + 1)) { if (++sp > (1u << 24)) __trap(); } } }
        __syncthreads();
        if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(bar), "r"((unsigned)(it + 1)) : "memory");
      }
      if (threadIdx.x == 0) { { unsigned sp = 0; while (ld_acq(bar) < (unsigned)(it + 1)) { if (++sp > (1u << 24)) __trap(); } } }
      __syncthreads();
    } else if (MODE == 4 || MODE == 5) {  // all-gather of per-CTA flags: one store, then every CTA polls all the flags (one hop, no atomics)
      const int stride = MODE == 4 ? 1 : 32;  // contiguous (4 lines for 128 CTAs) or one line per flag
      __syncthreads();
      if (threadIdx.x == 0) { epoch++; asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + blockIdx.x * stride), "r"(epoch) : "memory"); }
      if (threadIdx.x < gridDim.x) { unsigned sp = 0; while (ld_acq(flags + threadIdx.x * stride) < (unsigned)(it + 1)) { if (++sp > (1u << 24)) __trap(); } }
      __syncthreads();
    } else if (MODE == 6) {  // as 4, relaxed polling by one warp (4 flags per lane, 16-byte loads), one acquire fence at the end
      __syncthreads();
      if (threadIdx.x == 0) { epoch++; asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + blockIdx.x), "r"(epoch) : "memory"); }
      if (threadIdx.x < 32) {
        const unsigned want = (unsigned)(it + 1);
        unsigned sp = 0;
        while (true) {
          uint4 v;
          asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(flags + 4 * threadIdx.x) : "memory");
          const bool ok = (4 * threadIdx.x >= gridDim.x) || (v.x >= want && v.y >= want && v.z >= want && v.w >= want);
          if (__all_sync(0xffffffffu, ok)) break;
          if (++sp > (1u << 24)) __trap();
        }
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
      }
      __syncthreads();
    }
  }
}

template <int MODE>
float run(int grid, int csize, int iters) {
  unsigned *bar, *flags; cudaMalloc(&bar, 256); cudaMalloc(&flags, 128 * 256); cudaMemset(bar, 0, 256); cudaMemset(flags, 0, 128 * 256);
  cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(288); cfg.dynamicSmemBytes = 200 * 1024;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaLaunchKernelEx(&cfg, k<MODE>, bar, flags, 10, csize); cudaDeviceSynchronize(); cudaMemset(bar, 0, 256); cudaMemset(flags, 0, 128 * 256);
  cudaEventRecord(a); cudaLaunchKernelEx(&cfg, k<MODE>, bar, flags, iters, csize); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("err %s\n", cudaGetErrorString(e));
  cudaFree(bar); cudaFree(flags);
  return ms * 1e3f / iters;
}
int main() {
  const int iters = 2000;
  printf("flat red+acquire-spin      128 CTAs: %.2f us/barrier\n", run<0>(128, 4, iters));
  printf("hierarchical cluster(4)    128 CTAs: %.2f us/barrier\n", run<1>(128, 4, iters));
  printf("hierarchical cluster(2)    128 CTAs: %.2f us/barrier\n", run<1>(128, 2, iters));
  printf("flat relaxed-spin          128 CTAs: %.2f us/barrier\n", run<2>(128, 4, iters));
  printf("flags collected by CTA 0   128 CTAs: %.2f us/barrier\n", run<3>(128, 4, iters));
  printf("flat red+acquire-spin       32 CTAs: %.2f us/barrier\n", run<0>(32, 4, iters));
  printf("all-gather flags (packed)  128 CTAs: %.2f us/barrier\n", run<4>(128, 4, iters));
  printf("all-gather flags (padded)  128 CTAs: %.2f us/barrier\n", run<5>(128, 4, iters));
  printf("all-gather, one warp v4    128 CTAs: %.2f us/barrier\n", run<6>(128, 4, iters));
  return 0;
}
