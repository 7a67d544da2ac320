// interpolate.cu -- three_nn, three_interpolate (+grad) and their fusion for the FP modules.
//
// Reference kernels replaced (_ext-src/src/interpolate_gpu.cu): three_nn_kernel :9-59 (one CTA per scene, one thread
// per unknown point, `known` streamed from global by every thread), three_interpolate_kernel :72-101,
// three_interpolate_grad_kernel :116-143; weight computation in PointnetFPModule.forward
// (pointnet2_modules.py:382-385: dist = sqrt(dist2); w = 1/(dist+1e-8); w /= sum(w)).
//
// three_nn here: 128 unknown points per CTA (grid = ceil(n/128) x B), the known set staged through shared
// memory in 1024-point tiles.  The reference keeps its three running bests in double initialised to 1e40
// while every candidate distance is a float; comparing float values promoted to double is order-isomorphic
// to comparing the floats, and 1e40 (> FLT_MAX) behaves like +inf for every `d < best` test and converts to
// +inf on the final double->float store.  The float/+inf implementation below is therefore bit-identical.
#include "common.cuh"

namespace rfd {

constexpr int NN_THREADS = 128;
constexpr int NN_TILE = 1024;

__device__ __forceinline__ void nn3_update(float d, int k, float &b1, float &b2, float &b3, int &i1, int &i2,
                                           int &i3) {
  // reference :34-48, strict '<' chain: earliest index wins ties
  if (d < b1) {
    b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k;
  } else if (d < b2) {
    b3 = b2; i3 = i2; b2 = d; i2 = k;
  } else if (d < b3) {
    b3 = d; i3 = k;
  }
}

__device__ __forceinline__ void nn3_scan(const float *__restrict__ known, int m, float ux, float uy, float uz,
                                         bool active, float *s_known, float &b1, float &b2, float &b3, int &i1,
                                         int &i2, int &i3) {
  b1 = b2 = b3 = INFINITY;
  i1 = i2 = i3 = 0;
  for (int base = 0; base < m; base += NN_TILE) {
    const int cnt = min(NN_TILE, m - base);
    __syncthreads();
    for (int e = threadIdx.x; e < cnt * 3; e += blockDim.x) s_known[e] = __ldg(known + (size_t)base * 3 + e);
    __syncthreads();
    if (active) {
      for (int k = 0; k < cnt; ++k) {
        const float d = sqdist_yxz(ux - s_known[k * 3 + 0], uy - s_known[k * 3 + 1], uz - s_known[k * 3 + 2]);
        nn3_update(d, base + k, b1, b2, b3, i1, i2, i3);
      }
    }
  }
}

__global__ void __launch_bounds__(NN_THREADS)
three_nn_kernel(const float *__restrict__ unknown, const float *__restrict__ known, int n, int m,
                float *__restrict__ dist2, int *__restrict__ idx) {
  __shared__ float s_known[NN_TILE * 3];
  const int b = blockIdx.y;
  const int j = blockIdx.x * NN_THREADS + threadIdx.x;
  unknown += (size_t)b * n * 3;
  known += (size_t)b * m * 3;
  const bool active = j < n;
  float ux = 0.f, uy = 0.f, uz = 0.f;
  if (active) { ux = __ldg(unknown + j * 3 + 0); uy = __ldg(unknown + j * 3 + 1); uz = __ldg(unknown + j * 3 + 2); }
  float b1, b2, b3;
  int i1, i2, i3;
  nn3_scan(known, m, ux, uy, uz, active, s_known, b1, b2, b3, i1, i2, i3);
  if (active) {
    float *d = dist2 + ((size_t)b * n + j) * 3;
    int *ix = idx + ((size_t)b * n + j) * 3;
    d[0] = b1; d[1] = b2; d[2] = b3;
    ix[0] = i1; ix[1] = i2; ix[2] = i3;
  }
}

// out[b,c,j] = p[b,c,i1]*w1 + p[b,c,i2]*w2 + p[b,c,i3]*w3 in the contraction order of the sm_100 build of
// the reference (:98-99): t = p2*w2 ; t = fma(p1,w1,t) ; t = fma(p3,w3,t).
__device__ __forceinline__ float interp3(float p1, float p2, float p3, float w1, float w2, float w3) {
  float t = __fmul_rn(p2, w2);
  t = __fmaf_rn(p1, w1, t);
  return __fmaf_rn(p3, w3, t);
}

__global__ void __launch_bounds__(256)
three_interpolate_kernel(const float *__restrict__ points, const int *__restrict__ idx,
                         const float *__restrict__ weight, int c, int m, int n, float *__restrict__ out) {
  const int b = blockIdx.z;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int *ix = idx + ((size_t)b * n + j) * 3;
  const float *w = weight + ((size_t)b * n + j) * 3;
  const int i1 = __ldg(ix), i2 = __ldg(ix + 1), i3 = __ldg(ix + 2);
  const float w1 = __ldg(w), w2 = __ldg(w + 1), w3 = __ldg(w + 2);
  for (int l = blockIdx.y; l < c; l += gridDim.y) {
    const float *p = points + ((size_t)b * c + l) * m;
    out[((size_t)b * c + l) * n + j] = interp3(__ldg(p + i1), __ldg(p + i2), __ldg(p + i3), w1, w2, w3);
  }
}

__global__ void __launch_bounds__(256)
three_interpolate_grad_kernel(const float *__restrict__ grad_out, const int *__restrict__ idx,
                              const float *__restrict__ weight, int c, int n, int m, float *__restrict__ grad_points) {
  const int b = blockIdx.z;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int *ix = idx + ((size_t)b * n + j) * 3;
  const float *w = weight + ((size_t)b * n + j) * 3;
  const int i1 = __ldg(ix), i2 = __ldg(ix + 1), i3 = __ldg(ix + 2);
  const float w1 = __ldg(w), w2 = __ldg(w + 1), w3 = __ldg(w + 2);
  for (int l = blockIdx.y; l < c; l += gridDim.y) {
    const float g = __ldg(grad_out + ((size_t)b * c + l) * n + j);
    float *gp = grad_points + ((size_t)b * c + l) * m;
    atomicAdd(gp + i1, g * w1);  // reference :139-141
    atomicAdd(gp + i2, g * w2);
    atomicAdd(gp + i3, g * w3);
  }
}

// Fused FP front end: 3-NN search + weights + interpolation into out[b, 0:C, j] of a (B,Ctot,n) tensor.
// Weight arithmetic mirrors torch's elementwise ops in pointnet2_modules.py:382-385 in fp32:
//   dist = sqrt(dist2) (IEEE) ; r = 1.0f / (dist + 1e-8f) ; norm = r0 + r1 + r2 (torch.sum over 3 elements:
//   sequential order) ; w = r / norm.
__global__ void __launch_bounds__(NN_THREADS)
three_nn_interpolate_kernel(const float *__restrict__ unknown, const float *__restrict__ known,
                            const float *__restrict__ known_feats, int n, int m, int C, int Ctot,
                            float *__restrict__ out) {
  __shared__ float s_known[NN_TILE * 3];
  const int b = blockIdx.y;
  const int j = blockIdx.x * NN_THREADS + threadIdx.x;
  unknown += (size_t)b * n * 3;
  known += (size_t)b * m * 3;
  const bool active = j < n;
  float ux = 0.f, uy = 0.f, uz = 0.f;
  if (active) { ux = __ldg(unknown + j * 3 + 0); uy = __ldg(unknown + j * 3 + 1); uz = __ldg(unknown + j * 3 + 2); }
  float b1, b2, b3;
  int i1, i2, i3;
  nn3_scan(known, m, ux, uy, uz, active, s_known, b1, b2, b3, i1, i2, i3);
  if (!active) return;
  const float r1 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(b1), 1e-8f));
  const float r2 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(b2), 1e-8f));
  const float r3 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(b3), 1e-8f));
  const float norm = __fadd_rn(__fadd_rn(r1, r2), r3);
  const float w1 = __fdiv_rn(r1, norm), w2 = __fdiv_rn(r2, norm), w3 = __fdiv_rn(r3, norm);
  const float *f = known_feats + (size_t)b * C * m;
  float *o = out + (size_t)b * Ctot * n + j;
  // blockIdx.z owns a slice of the channels: the (cheap) neighbour search is repeated per slice so that the 3 C
  // dependent gathers of a point are spread over gridDim.z CTAs instead of one thread (FP1/FP2: 16 / 32 CTAs of 128
  // threads took 0.14 / 0.16 ms with one thread walking all 256 channels)
  const int cper = (C + gridDim.z - 1) / gridDim.z;
  const int l0 = blockIdx.z * cper, l1 = min(C, l0 + cper);
#pragma unroll 4
  for (int l = l0; l < l1; ++l) {
    const float *p = f + (size_t)l * m;
    o[(size_t)l * n] = interp3(__ldg(p + i1), __ldg(p + i2), __ldg(p + i3), w1, w2, w3);
  }
}

}  // namespace rfd

using namespace rfd;

extern "C" int rfd_three_nn(const float *unknown, const float *known, int B, int n, int m, float *dist2, int *idx,
                            void *stream) {
  if (B < 0 || n < 0 || m < 0) return RFD_ERR_INVALID_ARGUMENT;
  if (B == 0 || n == 0) return RFD_OK;
  if (!unknown || (m > 0 && !known) || !dist2 || !idx) return RFD_ERR_INVALID_ARGUMENT;
  if (B > 65535) return RFD_ERR_UNSUPPORTED_SIZE;
  dim3 grid(h_ceil_div(n, NN_THREADS), B);
  three_nn_kernel<<<grid, NN_THREADS, 0, as_stream(stream)>>>(unknown, known, n, m, dist2, idx);
  RFD_CHECK_LAUNCH("three_nn_kernel");
  return RFD_OK;
}

extern "C" int rfd_three_interpolate(const float *points, const int *idx, const float *weight, int B, int C, int m,
                                     int n, float *out, void *stream) {
  if (B < 0 || C < 0 || m < 0 || n < 0) return RFD_ERR_INVALID_ARGUMENT;
  if ((long long)B * C * n == 0) return RFD_OK;
  if (!points || !idx || !weight || !out) return RFD_ERR_INVALID_ARGUMENT;
  if (B > 65535) return RFD_ERR_UNSUPPORTED_SIZE;
  dim3 grid(h_ceil_div(n, 256), C < 64 ? C : 64, B);
  three_interpolate_kernel<<<grid, 256, 0, as_stream(stream)>>>(points, idx, weight, C, m, n, out);
  RFD_CHECK_LAUNCH("three_interpolate_kernel");
  return RFD_OK;
}

extern "C" int rfd_three_interpolate_grad(const float *grad_out, const int *idx, const float *weight, int B, int C,
                                          int n, int m, float *grad_points, void *stream) {
  if (B < 0 || C < 0 || m < 0 || n < 0) return RFD_ERR_INVALID_ARGUMENT;
  if ((long long)B * C * m == 0) return RFD_OK;
  if (!grad_points) return RFD_ERR_INVALID_ARGUMENT;
  RFD_CHECK_CUDA(cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)B * C * m, as_stream(stream)),
                 "three_interpolate_grad memset");
  if (n == 0) return RFD_OK;
  if (!grad_out || !idx || !weight) return RFD_ERR_INVALID_ARGUMENT;
  if (B > 65535) return RFD_ERR_UNSUPPORTED_SIZE;
  dim3 grid(h_ceil_div(n, 256), C < 64 ? C : 64, B);
  three_interpolate_grad_kernel<<<grid, 256, 0, as_stream(stream)>>>(grad_out, idx, weight, C, n, m, grad_points);
  RFD_CHECK_LAUNCH("three_interpolate_grad_kernel");
  return RFD_OK;
}

extern "C" int rfd_three_nn_interpolate(const float *unknown, const float *known, const float *known_feats, int B,
                                        int n, int m, int C, int Ctot, float *out, void *stream) {
  if (B < 0 || n < 0 || m < 1 || C < 0 || Ctot < C) return RFD_ERR_INVALID_ARGUMENT;
  if (B == 0 || n == 0 || C == 0) return RFD_OK;
  if (!unknown || !known || !known_feats || !out) return RFD_ERR_INVALID_ARGUMENT;
  if (B > 65535) return RFD_ERR_UNSUPPORTED_SIZE;
  // channel slices: enough CTAs to fill the GPU, at least 8 channels per slice
  int cs = 1;
  const long long base_ctas = (long long)h_ceil_div(n, NN_THREADS) * B;
  while (cs < 32 && base_ctas * cs < 592 && C / (cs * 2) >= 8) cs *= 2;
  dim3 grid(h_ceil_div(n, NN_THREADS), B, cs);
  three_nn_interpolate_kernel<<<grid, NN_THREADS, 0, as_stream(stream)>>>(unknown, known, known_feats, n, m, C, Ctot,
                                                                           out);
  RFD_CHECK_LAUNCH("three_nn_interpolate_kernel");
  return RFD_OK;
}
