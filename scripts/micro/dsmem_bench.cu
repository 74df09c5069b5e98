// All-to-all exchange inside a 16-CTA cluster (the h exchange of rec_tc.cu): every CTA sends 1 KB to each of the 16 CTAs per
// step (16 KB out, 16 KB in), receivers count transaction bytes on an mbarrier.  (a) st.async 16 B per thread, 64 B contiguous
// per row (what rec_tc does), (b) cp.async.bulk.shared::cluster of 512 B, (c) of 1024 B.  Reports us per step.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank)); return r; }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  do { asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory"); } while (!done);
}
__device__ __forceinline__ void st_async(uint32_t raddr, uint4 v, uint32_t rbar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(rbar) : "memory");
}
__device__ __forceinline__ void bulk_s2s(uint32_t rdst, uint32_t src, uint32_t bytes, uint32_t rbar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(rdst), "r"(src), "r"(bytes), "r"(rbar) : "memory");
}

constexpr int G = 16;
// mode 0: st.async; 1: bulk 512 B x 2 per destination; 2: bulk 1024 B x 1; groups = independent exchanges per step (NG)
__global__ void __cluster_dims__(G, 1, 1) __launch_bounds__(128, 1) k(int mode, int groups, int steps, long long* out) {
  extern __shared__ __align__(1024) unsigned char dyn[];
  unsigned char (*recv)[2][G * 1024] = reinterpret_cast<unsigned char (*)[2][G * 1024]>(dyn);  // [group][parity][sender]
  __shared__ __align__(128) unsigned char mine[2][1024];
  __shared__ __align__(8) unsigned long long bar[2][2];
  const int tid = threadIdx.x;
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (tid == 0) for (int g = 0; g < 2; ++g) for (int p = 0; p < 2; ++p) mbar_init(smem_u32(&bar[g][p]), 1);
  for (int i = tid; i < 2 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(mine)[i] = i;
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (tid == 0) for (int g = 0; g < groups; ++g) { mbar_expect(smem_u32(&bar[g][0]), G * 1024); mbar_expect(smem_u32(&bar[g][1]), G * 1024); }
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  const long long t0 = clock64();
  for (int s = 0; s < steps; ++s) {
    const int par = s & 1;
    for (int g = 0; g < groups; ++g) {
      const uint32_t dst_local = smem_u32(&recv[g][par][rank * 1024]);
      const uint32_t bar_local = smem_u32(&bar[g][par]);
      if (mode == 0) {
        for (int j = 0; j < G * 64 / 128; ++j) {  // 16 dest x 64 chunks of 16 B
          const int idx = tid + j * 128;
          const int dst = idx >> 6, ch = idx & 63;
          const uint4 v = *reinterpret_cast<const uint4*>(&mine[g][ch * 16]);
          st_async(mapa(dst_local + ch * 16, dst), v, mapa(bar_local, dst));
        }
      } else {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const int piece = mode == 1 ? 512 : 1024, per = 1024 / piece;
        if (tid < G * per) {
          const int dst = tid / per, o = (tid % per) * piece;
          bulk_s2s(mapa(dst_local + o, dst), smem_u32(&mine[g][o]), piece, mapa(bar_local, dst));
        }
      }
    }
    for (int g = 0; g < groups; ++g) {
      if (tid == 0) {
        mbar_wait(smem_u32(&bar[g][par]), (s >> 1) & 1);
        if (s + 2 < steps) mbar_expect(smem_u32(&bar[g][par]), G * 1024);
      }
    }
    __syncthreads();
  }
  const long long t1 = clock64();
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (blockIdx.x == 0 && tid == 0) out[0] = t1 - t0;
}

int main() {
  long long* out; cudaMalloc(&out, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 2 * G * 1024);
  const int steps = 2000;
  const char* names[3] = {"st.async 16B", "bulk 512B", "bulk 1024B"};
  for (int groups = 1; groups <= 2; ++groups)
    for (int mode = 0; mode < 3; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        k<<<4 * G, 128, 2 * 2 * G * 1024>>>(mode, groups, steps, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      }
      long long c; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
      printf("%-14s groups %d: %.3f us per step (%.1f B/clk per SM in+out)\n", names[mode], groups, c / 1.965e3 / steps, groups * 2.0 * G * 1024 * steps / c);
    }
  return 0;
}
