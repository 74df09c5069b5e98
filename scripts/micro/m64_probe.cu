// Where does tcgen05.mma cta_group::1 kind::f16 with M = 64 put row i of D in tensor memory?  A[i][0] = i + 1, B[n][0] = 1
// -> D[i][n] = i + 1; TMEM is pre-filled with -1; every lane's 16 columns are dumped.
#include <cstdio>
#include "../../phones_las_b200/csrc/common.cuh"
#include "../../phones_las_b200/csrc/tcgen05.cuh"
using namespace plas;

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
               "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}

__global__ void __launch_bounds__(128, 1) probe(float* out, int M) {
  extern __shared__ unsigned char raw_[];
  const uint32_t raw = smem_u32(raw_);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* smem = raw_ + (base - raw);
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) unsigned long long s_bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  // A tile at +0 (128 rows x 128 B, SW128), B tile at +16384 (16 rows)
  for (int i = tid; i < (16384 + 2048) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  __syncthreads();
  if (tid < 128) {  // row tid of A: element k = 0 sits in chunk 0 ^ (r & 7)
    __nv_bfloat16* a = reinterpret_cast<__nv_bfloat16*>(smem + tid * 128 + (((0 ^ (tid & 7))) << 4));
    a[0] = __float2bfloat16((float)(tid + 1));
  }
  if (tid < 16) {
    __nv_bfloat16* b = reinterpret_cast<__nv_bfloat16*>(smem + 16384 + tid * 128 + (((0 ^ (tid & 7))) << 4));
    b[0] = __float2bfloat16(1.0f);
  }
  if (tid == 0) mbar_init(smem_u32(&s_bar), 1);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = s_tmem;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  uint32_t fill[16];
  for (int i = 0; i < 16; ++i) fill[i] = __float_as_uint(-1.0f);
  tmem_st16(tb + lane_base, fill);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    if (elect_one()) {
      tc_fence_after();
      umma_bf16(tb, umma_smem_desc(base), umma_smem_desc(base + 16384), umma_idesc_bf16(M, 16), 0u);
      umma_commit(smem_u32(&s_bar));
    }
    __syncwarp();
  }
  if (tid % 32 == 0) mbar_wait(smem_u32(&s_bar), 0);
  __syncwarp();
  tc_fence_after();
  uint32_t r[16];
  tmem_ld16(tb + lane_base, r);
  tmem_ld_wait();
  for (int i = 0; i < 16; ++i) out[tid * 16 + i] = __uint_as_float(r[i]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tb) : "memory");
}

int main() {
  float* out; cudaMalloc(&out, 128 * 16 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  for (int M : {128, 64}) {
    probe<<<1, 128, 32768>>>(out, M);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("M=%d error %s\n", M, cudaGetErrorString(e)); return 1; }
    float h[128 * 16]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("M=%d: lane -> D row (col 0 | col 8 | col 15)\n", M);
    for (int l = 0; l < 128; ++l) printf("%s%d:%g|%g|%g", (l % 8) ? "  " : "\n", l, h[l * 16], h[l * 16 + 8], h[l * 16 + 15]);
    printf("\n");
  }
  return 0;
}
