// How fast do 8 x 8 KB bulk copies (cp.async.bulk global -> shared) land in one SM when (a) every CTA reads the SAME 64 KB,
// (b) every CTA reads its own 64 KB, and how does a cooperative LDG.128 + STS copy of the same bytes compare?  128 CTAs.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// mode 0: bulk, shared source; 1: bulk, private source; 2: LDG+STS shared source; 3: LDG+STS private; piece = bytes per bulk copy
__global__ void __launch_bounds__(256, 1) k(const unsigned char* src, int mode, int piece, int reps, unsigned long long* out, unsigned* gbar) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ __align__(8) unsigned long long bars[8];
  const int tid = threadIdx.x;
  if (tid == 0) for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1);
  __syncthreads();
  const unsigned char* my = src + ((mode & 1) ? (size_t)(blockIdx.x + 1) * 65536 : 0);
  unsigned long long tsum[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  uint32_t parity = 0;
  for (int r = 0; r < reps; ++r) {
    if (mode == 4 || mode == 5) {  // every CTA rewrites its own 8 bytes of every row of the shared tiles (like the decoder's epilogue)
      if (tid < 64) {
        const int u = blockIdx.x * 4;
        unsigned char* w = const_cast<unsigned char*>(src) + (size_t)(u >> 6) * 8192 + tid * 128 + ((((u & 63) >> 3) ^ (tid & 7)) << 4) + ((u & 7) << 1);
        *reinterpret_cast<uint2*>(w) = make_uint2(r, blockIdx.x);
      }
      if (mode == 5) __threadfence();
      asm volatile("fence.proxy.async;" ::: "memory");
    }
    // grid barrier so that all CTAs start together (as after the decoder's phase barrier)
    __syncthreads();
    if (tid == 0) {
      asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(gbar), "r"(1u) : "memory");
      unsigned v;
      do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(gbar) : "memory"); } while (v < (unsigned)(r + 1) * gridDim.x);
    }
    __syncthreads();
    const unsigned long long t0 = gtime();
    if (mode < 2 || mode >= 4) {
      if (mode >= 4) asm volatile("fence.proxy.async;" ::: "memory");
      if (tid == 0) {
        const int rot = blockIdx.x % 8;
        for (int i = 0; i < 8; ++i) {
          const int kb = (rot + i) % 8;
          mbar_expect(smem_u32(&bars[i]), 8192);
          for (int o = 0; o < 8192; o += piece) bulk(smem_u32(sm) + i * 8192 + o, my + kb * 8192 + o, piece, smem_u32(&bars[i]));
        }
        for (int i = 0; i < 8; ++i) {
          mbar_wait(smem_u32(&bars[i]), parity);
          tsum[i] += gtime() - t0;
        }
      }
      parity ^= 1u;
    } else {
      uint4 v[16];
      const uint4* s4 = reinterpret_cast<const uint4*>(my);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = __ldcg(s4 + i * 256 + tid);
#pragma unroll
      for (int i = 0; i < 16; ++i) reinterpret_cast<uint4*>(sm)[i * 256 + tid] = v[i];
      __syncthreads();
      if (tid == 0) tsum[7] += gtime() - t0;
    }
    __syncthreads();
    tsum[8] += gtime() - t0;
  }
  if (blockIdx.x == 0 && tid == 0) for (int i = 0; i < 9; ++i) out[i] = tsum[i];
}

int main() {
  unsigned char* src; unsigned long long* out; unsigned* gbar;
  cudaMalloc(&src, 129 * 65536); cudaMemset(src, 1, 129 * 65536);
  cudaMalloc(&out, 9 * 8); cudaMalloc(&gbar, 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024);
  const int reps = 200;
  const char* names[6] = {"bulk shared", "bulk private", "ldg shared", "ldg private", "bulk fresh", "bulk fresh+tf"};
  for (int piece : {8192}) {
    for (int mode = 0; mode < 6; ++mode) {
      if (mode >= 2 && piece != 8192) continue;
      for (int pass = 0; pass < 2; ++pass) {
        cudaMemset(gbar, 0, 4);
        void* args[] = {&src, &mode, &piece, (void*)&reps, &out, &gbar};
        cudaLaunchCooperativeKernel((const void*)k, dim3(128), dim3(256), args, 65536 + 1024, 0);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      }
      unsigned long long h[9]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      printf("%-13s piece %5d: tile arrival (us)", names[mode], piece);
      for (int i = 0; i < 8; ++i) printf(" %.2f", h[i] / 1e3 / reps);
      printf("  | all done %.2f\n", h[8] / 1e3 / reps);
    }
  }
  return 0;
}
