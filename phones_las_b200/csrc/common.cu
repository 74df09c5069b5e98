// Error plumbing and small layout kernels (cast/pad, time masking) of libplas.so.
#include <stdarg.h>
#include <mutex>

#include "common.cuh"
#include "../../include/plas.h"

namespace plas {

char* err_buf() {
  static thread_local char buf[1024] = {0};
  return buf;
}

int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 1024, fmt, ap);
  va_end(ap);
  return code;
}

int num_sms() {
  static int sms = 0;
  static std::once_flag once;
  std::call_once(once, [] {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  });
  return sms;
}

__global__ void cast_pad_bf16_kernel(const float* __restrict__ x, long long rows, int cols,
                                     long long ld_in, __nv_bfloat16* __restrict__ y, long long ld_out) {
  const long long total = rows * ld_out;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / ld_out;
    const int c = (int)(i - r * ld_out);
    y[i] = __float2bfloat16_rn(c < cols ? x[r * ld_in + c] : 0.f);
  }
}

template <typename T>
__global__ void mask_time_kernel(const T* __restrict__ x, T* __restrict__ y, const int* __restrict__ len,
                                 int B, int T_, int D) {
  const long long total = (long long)B * T_ * D;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long bt = i / D;
    const int t = (int)(bt % T_);
    const int b = (int)(bt / T_);
    y[i] = t < len[b] ? x[i] : from_f32<T>(0.f);
  }
}

}  // namespace plas

using namespace plas;

extern "C" const char* plas_last_error(void) { return err_buf(); }
extern "C" int plas_version(void) { return 100; }
extern "C" int plas_num_sms(void) { return num_sms(); }

static int grid_for(long long total) {
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)(num_sms() > 0 ? num_sms() : 148) * 16;
  return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

extern "C" int plas_cast_pad_bf16(const float* x, int64_t rows, int32_t cols, int64_t ld_in, void* y,
                                  int64_t ld_out, plas_stream_t stream) {
  PLAS_REQUIRE(x && y && rows >= 0 && cols > 0 && ld_in >= cols && ld_out >= cols, "cast_pad: bad shape");
  if (rows == 0) return PLAS_OK;
  cast_pad_bf16_kernel<<<grid_for(rows * ld_out), 256, 0, (cudaStream_t)stream>>>(
      x, rows, cols, ld_in, (__nv_bfloat16*)y, ld_out);
  PLAS_CUDA(cudaGetLastError());
  return PLAS_OK;
}

extern "C" int plas_mask_time(int32_t dtype, const void* x, void* y, const int32_t* len, int32_t B,
                              int32_t T, int32_t D, plas_stream_t stream) {
  PLAS_REQUIRE(x && y && len && B > 0 && T > 0 && D > 0, "mask_time: bad argument");
  const long long total = (long long)B * T * D;
  if (dtype == PLAS_F32)
    mask_time_kernel<float><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const float*)x, (float*)y, len, B, T, D);
  else if (dtype == PLAS_BF16)
    mask_time_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, (__nv_bfloat16*)y, len, B, T, D);
  else
    return set_err(PLAS_EINVAL, "mask_time: dtype %d", dtype);
  PLAS_CUDA(cudaGetLastError());
  return PLAS_OK;
}
