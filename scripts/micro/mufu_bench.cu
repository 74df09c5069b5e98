// MUFU throughput probe (sm_100a): lanes per clock per SM of tanh.approx / ex2.approx / rcp.approx, and of an
// FMA-pipe rational tanh, with 8 warps per SM (the decoder's occupancy) and 8 independent chains per thread.
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__device__ __forceinline__ float op(float x) {
  float y;
  if (OP == 0) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  else if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  else if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  else {  // odd rational on the FMA pipe + one rcp
    const float x2 = x * x;
    float p = fmaf(x2, 2.0e-5f, 1.0e-3f); p = fmaf(p, x2, 5.0e-2f); p = fmaf(p, x2, 1.0f);
    float q = fmaf(x2, 1.0e-4f, 1.0e-2f); q = fmaf(q, x2, 4.0e-1f); q = fmaf(q, x2, 1.0f);
    float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(q));
    y = x * p * r;
  }
  return y;
}

template <int OP>
__global__ void __launch_bounds__(256, 1) k(float* out, int iters, long long* cyc) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0.001f * (threadIdx.x + i);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = op<OP>(v[i] + 0.25f);
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 4096;
  k<OP><<<148, 256>>>(out, iters, cyc);
  k<OP><<<148, 256>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = h[0];
  printf("%-10s %.2f lanes/clk/SM (cycles %.0f for %d x 8 x 256 ops)\n", name, (double)iters * 8 * 256 / c, c, iters);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("tanh"); run<1>("ex2"); run<2>("rcp"); run<3>("rational");
  return 0;
}
