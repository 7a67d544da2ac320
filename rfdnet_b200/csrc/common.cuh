// common.cuh -- shared helpers for librfdnet_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>

#include "../../include/rfdnet_b200.h"

namespace rfd {

extern thread_local char g_last_error[512];
extern std::atomic<long long> g_launch_count;

int set_cuda_error(cudaError_t e, const char *where);

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// call after every kernel launch
#define RFD_CHECK_LAUNCH(where)                                  \
  do {                                                           \
    ::rfd::g_launch_count.fetch_add(1, std::memory_order_relaxed); \
    cudaError_t _e = cudaGetLastError();                         \
    if (_e != cudaSuccess) return ::rfd::set_cuda_error(_e, where); \
  } while (0)

#define RFD_CHECK_CUDA(expr, where)                               \
  do {                                                            \
    cudaError_t _e = (expr);                                      \
    if (_e != cudaSuccess) return ::rfd::set_cuda_error(_e, where); \
  } while (0)

// Squared distance in the exact operation order of the reference kernels as compiled by nvcc 12.9
// for sm_100 (cuobjdump -sass): t = dy*dy ; t = fma(dx,dx,t) ; d = fma(dz,dz,t).
__device__ __forceinline__ float sqdist_yxz(float dx, float dy, float dz) {
  float t = __fmul_rn(dy, dy);
  t = __fmaf_rn(dx, dx, t);
  return __fmaf_rn(dz, dz, t);
}

__device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int h_ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// fp32 pointwise layer (mlp_f32.cu); pre_scale/pre_shift (B x pre_bstride) optional input affine+ReLU
int launch_pointwise_f32(const float *x, const float *W, const float *scale, const float *shift,
                         const float *residual, const float *pre_scale, const float *pre_shift,
                         long long pre_bstride, int relu, int pool, int B, int Cin, int Cout, int L, float *y,
                         cudaStream_t stream);

}  // namespace rfd
