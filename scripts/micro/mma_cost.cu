// What does one small tcgen05.mma (kind::f16, K = 16) cost as a function of the operand form and the tile shape?  A chain of
// NM MMAs into one accumulator, committed and waited for, repeated; the slope between NM = 32 and NM = 64 is the per-MMA cost.
//   SS: A [M x 16] and B [N x 16] both from shared memory (K-major SWIZZLE_128B tiles);  TS: A from tensor memory.
#include <cstdio>
#include "../../phones_las_b200/csrc/common.cuh"
#include "../../phones_las_b200/csrc/tcgen05.cuh"
using namespace plas;

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(128, 1) cost(long long* out, int M, int N, int ts, int NM, int reps) {
  extern __shared__ unsigned char raw_[];
  const uint32_t raw = smem_u32(raw_);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* smem = raw_ + (base - raw);
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) unsigned long long s_bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t A0 = base, B0 = base + 8 * 16384;  // 8 k blocks of A (128 rows x 128 B), then 8 k blocks of B (N rows x 128 B)
  for (int i = tid; i < (8 * 16384 + 8 * N * 128) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (tid == 0) mbar_init(smem_u32(&s_bar), 1);
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
      const uint32_t idesc = umma_idesc_bf16(M, N);
      uint32_t par = 0;
      long long t0 = 0;
      for (int rep = -2; rep < reps; ++rep) {
        if (rep == 0) t0 = clock64();
        const uint64_t a0 = umma_smem_desc(A0), b0 = umma_smem_desc(B0);
        const uint32_t bstep = (uint32_t)(N * 128) >> 4;
        for (int pass = 0; pass < NM / 32; ++pass) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {  // descriptors are base + compile-time offsets: nothing but the MMA itself per iteration
            const int kb = i >> 2, k = i & 3;
            const uint64_t bdesc = b0 + (uint64_t)(kb * bstep + 2 * k);
            if (ts) umma_ts(tb, tb + 128 + 8 * i, bdesc, idesc, (pass | i) != 0);
            else umma_bf16(tb, a0 + (uint64_t)(kb * 1024 + 2 * k), bdesc, idesc, (pass | i) != 0);
          }
        }
        umma_commit(smem_u32(&s_bar));
        mbar_wait(smem_u32(&s_bar), par);
        par ^= 1u;
        tc_fence_after();
      }
      out[blockIdx.x] = clock64() - t0;
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

int main() {
  long long* out; cudaMalloc(&out, 148 * 8);
  const int smem = 1024 + 8 * 16384 + 8 * 64 * 128;
  cudaFuncSetAttribute(cost, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 200;
  printf("form  M   N   cycles/chain(32)  cycles/chain(64)  cycles per MMA (slope)\n");
  for (int ts = 0; ts < 2; ++ts)
    for (int M : {64, 128})
      for (int N : {16, 32, 48, 64}) {
        if (ts && M == 64) continue;
        double c[2];
        for (int j = 0; j < 2; ++j) {
          cost<<<148, 128, smem>>>(out, M, N, ts, 32 << j, reps);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
          double s = 0; for (int i = 0; i < 148; ++i) s += h[i];
          c[j] = s / 148 / reps;
        }
        printf("%s  %3d %3d   %10.0f        %10.0f        %8.1f\n", ts ? "TS" : "SS", M, N, c[0], c[1], (c[1] - c[0]) / 32);
      }
  return 0;
}
