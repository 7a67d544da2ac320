// ballquery_group.cu -- ball query, grouping gather, their fusion, point gather, and the scatter-add grads.
//
// Reference kernels replaced (paths relative to _ext-src/src/):
//   query_ball_point_kernel   ball_query_gpu.cu:9-44     one CTA per scene, one THREAD per query scanning all N
//   group_points_kernel       group_points_gpu.cu:8-28   one CTA per scene, 4-byte random gathers
//   group_points_grad_kernel  group_points_gpu.cu:43-64
//   gather_points_kernel      sampling_gpu.cu:8-20 ; gather_points_grad_kernel :34-47
//   QueryAndGroup.forward     pointnet2_utils.py:302-361  (ball_query + group(xyz) + sub + div + group(feat) + cat)
//
// Design: one WARP per query.  The 32 lanes test 32 consecutive candidate points per step (coalesced reads of
// the shared point stream, which stays in L1/L2), a ballot + popc prefix assigns output slots in ascending index
// order -- exactly the "first nsample hits by ascending k" semantics of the reference -- and the warp stops as
// soon as nsample hits are found.  In the fused kernel the 8 warps of a CTA own 8 consecutive queries; after the
// scan the CTA writes the (3+C, 8*S) output tile channel by channel with fully coalesced 128-byte stores, the
// normalised relative xyz ((p - c) * (1/r)) computed on the fly, so none of the reference's four intermediate
// passes over the (B,3+C,M,S) tensor exists.
#include "common.cuh"

namespace rfd {

constexpr int BQ_WARPS = 8;  // queries per CTA
constexpr int BQ_THREADS = BQ_WARPS * 32;
constexpr int QG_MAX_S = 128;

// Scan for one query by one warp.  Hits are reported through `emit(pos, k)` in ascending k, pos < nsample.
// Returns the hit count (capped at nsample) and the first hit index.
template <typename Emit>
__device__ __forceinline__ int ball_scan(const float *__restrict__ xyz, int n, float qx, float qy, float qz,
                                         float radius2, int nsample, int lane, int &first, Emit emit) {
  int cnt = 0;
  first = 0;
  for (int base = 0; base < n && cnt < nsample; base += 32) {
    const int k = base + lane;
    bool hit = false;
    if (k < n) {
      const float x = __ldg(xyz + (size_t)k * 3 + 0);
      const float y = __ldg(xyz + (size_t)k * 3 + 1);
      const float z = __ldg(xyz + (size_t)k * 3 + 2);
      // reference :31-34: d2 = (new_x-x)^2 + (new_y-y)^2 + (new_z-z)^2 ; hit iff d2 < radius2 (NaN -> no hit)
      const float d2 = sqdist_yxz(qx - x, qy - y, qz - z);
      hit = d2 < radius2;
    }
    const unsigned mask = __ballot_sync(0xffffffffu, hit);
    if (mask) {
      if (cnt == 0) first = base + __ffs(mask) - 1;
      const int pos = cnt + __popc(mask & ((1u << lane) - 1u));
      if (hit && pos < nsample) emit(pos, k);
      cnt += __popc(mask);
    }
  }
  return cnt < nsample ? cnt : nsample;
}

__global__ void __launch_bounds__(BQ_THREADS)
ball_query_kernel(const float *__restrict__ new_xyz, const float *__restrict__ xyz, int n, int m, float radius,
                  int nsample, int *__restrict__ idx) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int j = blockIdx.x * BQ_WARPS + warp;
  if (j >= m) return;
  xyz += (size_t)b * n * 3;
  new_xyz += (size_t)b * m * 3;
  int *row = idx + ((size_t)b * m + j) * nsample;
  const float radius2 = __fmul_rn(radius, radius);  // reference :22
  const float qx = __ldg(new_xyz + j * 3 + 0), qy = __ldg(new_xyz + j * 3 + 1), qz = __ldg(new_xyz + j * 3 + 2);
  int first;
  const int cnt = ball_scan(xyz, n, qx, qy, qz, radius2, nsample, lane, first,
                            [&](int pos, int k) { row[pos] = k; });
  // reference :35-39: the first hit pre-fills every slot; no hit at all leaves the zero-initialised row
  const int fill = cnt == 0 ? 0 : first;
  for (int l = cnt + lane; l < nsample; l += 32) row[l] = fill;
}

// fused ball query + group.  grid = (ceil(M/8), B)
__global__ void __launch_bounds__(BQ_THREADS)
query_and_group_kernel(const float *__restrict__ xyz, const float *__restrict__ new_xyz,
                       const float *__restrict__ features, int n, int m, int C, float radius, int nsample,
                       int use_xyz, int normalize_xyz, float *__restrict__ new_features,
                       float *__restrict__ grouped_xyz, int *__restrict__ idx_out) {
  __shared__ int s_idx[BQ_WARPS][QG_MAX_S];
  __shared__ float s_q[BQ_WARPS][3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int j0 = blockIdx.x * BQ_WARPS;
  const int j = j0 + warp;
  xyz += (size_t)b * n * 3;
  new_xyz += (size_t)b * m * 3;
  const float radius2 = __fmul_rn(radius, radius);
  if (j < m) {
    const float qx = __ldg(new_xyz + j * 3 + 0), qy = __ldg(new_xyz + j * 3 + 1), qz = __ldg(new_xyz + j * 3 + 2);
    if (lane == 0) { s_q[warp][0] = qx; s_q[warp][1] = qy; s_q[warp][2] = qz; }
    int first;
    int *row = s_idx[warp];
    const int cnt = ball_scan(xyz, n, qx, qy, qz, radius2, nsample, lane, first,
                              [&](int pos, int k) { row[pos] = k; });
    const int fill = cnt == 0 ? 0 : first;
    for (int l = cnt + lane; l < nsample; l += 32) row[l] = fill;
  }
  __syncthreads();
  const int nq = min(BQ_WARPS, m - j0);  // queries of this CTA
  const int inner = nq * nsample;        // contiguous floats per channel: out[b][c][j0 .. j0+nq)[0..S)
  const size_t MS = (size_t)m * nsample;
  if (idx_out) {
    int *dst = idx_out + ((size_t)b * m + j0) * nsample;
    for (int e = threadIdx.x; e < inner; e += BQ_THREADS) dst[e] = s_idx[e / nsample][e % nsample];
  }
  const int cx = use_xyz ? 3 : 0;
  const int Ct = cx + C;
  // torch lowers `tensor /= python_float` on CUDA to a multiply with the f32 reciprocal (pointnet2_utils.py:337)
  const float inv_r = normalize_xyz ? __frcp_rn(radius) : 1.0f;
  // ---- relative xyz channels
  if (use_xyz || grouped_xyz) {
    for (int e = threadIdx.x; e < 3 * inner; e += BQ_THREADS) {
      const int c = e / inner, rem = e - c * inner;
      const int q = rem / nsample, s = rem - q * nsample;
      const int k = s_idx[q][s];
      float v = __fsub_rn(__ldg(xyz + (size_t)k * 3 + c), s_q[q][c]);  // :335 grouped_xyz -= new_xyz
      if (normalize_xyz) v = __fmul_rn(v, inv_r);                       // :337
      const size_t off = (size_t)c * MS + (size_t)j0 * nsample + rem;
      if (use_xyz) new_features[(size_t)b * Ct * MS + off] = v;
      if (grouped_xyz) grouped_xyz[(size_t)b * 3 * MS + off] = v;
    }
  }
  // ---- feature channels: out[b][cx+c][j][s] = features[b][c][idx]
  if (C > 0) {
    const float *__restrict__ f = features + (size_t)b * C * n;
    float *__restrict__ o = new_features + ((size_t)b * Ct + cx) * MS + (size_t)j0 * nsample;
    const int total = C * inner;
    for (int e = threadIdx.x; e < total; e += BQ_THREADS) {
      const int c = e / inner, rem = e - c * inner;
      const int q = rem / nsample, s = rem - q * nsample;
      o[(size_t)c * MS + rem] = __ldg(f + (size_t)c * n + s_idx[q][s]);
    }
  }
}

// group_points: out[b,c,j,s] = points[b,c,idx[b,j,s]].  One thread per output element, coalesced along (j,s).
__global__ void __launch_bounds__(256)
group_points_kernel(const float *__restrict__ points, const int *__restrict__ idx, int c, int n, long long ms,
                    float *__restrict__ out) {
  const int b = blockIdx.z, l = blockIdx.y;
  const float *__restrict__ p = points + ((size_t)b * c + l) * n;
  const int *__restrict__ ix = idx + (size_t)b * ms;
  float *__restrict__ o = out + ((size_t)b * c + l) * ms;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < ms; e += (long long)gridDim.x * blockDim.x)
    o[e] = __ldg(p + __ldg(ix + e));
}

__global__ void __launch_bounds__(256)
group_points_grad_kernel(const float *__restrict__ grad_out, const int *__restrict__ idx, int c, int n, long long ms,
                         float *__restrict__ grad_points) {
  const int b = blockIdx.z, l = blockIdx.y;
  float *__restrict__ gp = grad_points + ((size_t)b * c + l) * n;
  const int *__restrict__ ix = idx + (size_t)b * ms;
  const float *__restrict__ g = grad_out + ((size_t)b * c + l) * ms;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < ms; e += (long long)gridDim.x * blockDim.x)
    atomicAdd(gp + __ldg(ix + e), __ldg(g + e));
}

static int launch_group(bool grad, const float *src, const int *idx, int B, int C, int N, long long MS, float *dst,
                        cudaStream_t st) {
  if (B == 0 || C == 0 || MS == 0) return RFD_OK;
  if (C > 65535 || B > 65535) return RFD_ERR_UNSUPPORTED_SIZE;
  int gx = (int)((MS + 255) / 256);
  if (gx > 1024) gx = 1024;
  dim3 grid(gx, C, B);
  if (grad)
    group_points_grad_kernel<<<grid, 256, 0, st>>>(src, idx, C, N, MS, dst);
  else
    group_points_kernel<<<grid, 256, 0, st>>>(src, idx, C, N, MS, dst);
  RFD_CHECK_LAUNCH(grad ? "group_points_grad_kernel" : "group_points_kernel");
  return RFD_OK;
}

}  // namespace rfd

using namespace rfd;

extern "C" int rfd_ball_query(const float *new_xyz, const float *xyz, int B, int N, int M, float radius, int nsample,
                              int *idx, void *stream) {
  if (B < 0 || N < 0 || M < 0 || nsample < 0) return RFD_ERR_INVALID_ARGUMENT;
  if (B == 0 || M == 0 || nsample == 0) return RFD_OK;
  if (!new_xyz || !xyz || !idx) return RFD_ERR_INVALID_ARGUMENT;
  if (B > 65535) return RFD_ERR_UNSUPPORTED_SIZE;
  dim3 grid(h_ceil_div(M, BQ_WARPS), B);
  ball_query_kernel<<<grid, BQ_THREADS, 0, as_stream(stream)>>>(new_xyz, xyz, N, M, radius, nsample, idx);
  RFD_CHECK_LAUNCH("ball_query_kernel");
  return RFD_OK;
}

extern "C" int rfd_query_and_group(const float *xyz, const float *new_xyz, const float *features, int B, int N, int M,
                                   int C, float radius, int nsample, int use_xyz, int normalize_xyz,
                                   float *new_features, float *grouped_xyz, int *idx, void *stream) {
  if (B < 0 || N < 0 || M < 0 || C < 0 || nsample < 0) return RFD_ERR_INVALID_ARGUMENT;
  if (B == 0 || M == 0 || nsample == 0) return RFD_OK;
  if (!xyz || !new_xyz || (C > 0 && !features) || !new_features) return RFD_ERR_INVALID_ARGUMENT;
  if (!use_xyz && C == 0) return RFD_ERR_INVALID_ARGUMENT;  // pointnet2_utils.py:347-350 assert
  if (nsample > QG_MAX_S || B > 65535) return RFD_ERR_UNSUPPORTED_SIZE;
  dim3 grid(h_ceil_div(M, BQ_WARPS), B);
  query_and_group_kernel<<<grid, BQ_THREADS, 0, as_stream(stream)>>>(xyz, new_xyz, features, N, M, C, radius,
                                                                      nsample, use_xyz, normalize_xyz, new_features,
                                                                      grouped_xyz, idx);
  RFD_CHECK_LAUNCH("query_and_group_kernel");
  return RFD_OK;
}

extern "C" int rfd_group_points(const float *points, const int *idx, int B, int C, int N, int M, int S, float *out,
                                void *stream) {
  if (B < 0 || C < 0 || N < 0 || M < 0 || S < 0) return RFD_ERR_INVALID_ARGUMENT;
  if ((long long)B * C * M * S == 0) return RFD_OK;
  if (!points || !idx || !out) return RFD_ERR_INVALID_ARGUMENT;
  return launch_group(false, points, idx, B, C, N, (long long)M * S, out, as_stream(stream));
}

extern "C" int rfd_group_points_grad(const float *grad_out, const int *idx, int B, int C, int N, int M, int S,
                                     float *grad_points, void *stream) {
  if (B < 0 || C < 0 || N < 0 || M < 0 || S < 0) return RFD_ERR_INVALID_ARGUMENT;
  if ((long long)B * C * N == 0) return RFD_OK;
  if (!grad_points) return RFD_ERR_INVALID_ARGUMENT;
  RFD_CHECK_CUDA(cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)B * C * N, as_stream(stream)),
                 "group_points_grad memset");
  if ((long long)M * S == 0) return RFD_OK;
  if (!grad_out || !idx) return RFD_ERR_INVALID_ARGUMENT;
  return launch_group(true, grad_out, idx, B, C, N, (long long)M * S, grad_points, as_stream(stream));
}

// gather_points is group_points with S = 1 (idx (B,M)); same for the grad.
extern "C" int rfd_gather_points(const float *points, const int *idx, int B, int C, int N, int M, float *out,
                                 void *stream) {
  return rfd_group_points(points, idx, B, C, N, M, 1, out, stream);
}

extern "C" int rfd_gather_points_grad(const float *grad_out, const int *idx, int B, int C, int N, int M,
                                      float *grad_points, void *stream) {
  return rfd_group_points_grad(grad_out, idx, B, C, N, M, 1, grad_points, stream);
}
