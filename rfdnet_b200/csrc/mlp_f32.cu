// mlp_f32.cu -- exact-fp32 pointwise (1x1 conv) layer with folded BatchNorm, ReLU, residual and max-pool.
//
// Replaces the cuDNN/cuBLAS calls behind nn.Conv2d(1x1)+BatchNorm2d+ReLU (+F.max_pool2d over nsample) of
// build_shared_mlp (pointnet2_modules.py:9-19, 237-243), the Conv1d+BN1d stacks of VotingModule
// (vote_module.py:46-48) and ProposalModule (proposal_module.py:112-114) and the FP-module MLPs
// (pointnet2_modules.py:395-405) in EVAL mode (running statistics => per-channel affine).
// This is the fp32-exact path (config 2 "fp32", tolerance 1e-4); the bf16 tensor-core path lives in
// sa_mlp_tc.cu / onet_decoder.cu.
//
// GEMM view per scene b:  Y[b] (Cout x L) = W (Cout x Cin) . X[b] (Cin x L)   (channel-major, L contiguous).
// CTA tile 64 (out channels) x 64 (positions), K step 16, 256 threads x (4x4) register tile, FFMA only.
#include "common.cuh"

namespace rfd {

constexpr int MT = 64, NT = 64, KT = 16;

__global__ void __launch_bounds__(256)
pointwise_mlp_f32_kernel(const float *__restrict__ x, const float *__restrict__ W, const float *__restrict__ scale,
                         const float *__restrict__ shift, const float *__restrict__ residual,
                         const float *__restrict__ pre_scale, const float *__restrict__ pre_shift, long long pre_bstride,
                         int relu, int pool, int Cin, int Cout, int L, float *__restrict__ y) {
  __shared__ float Ws[KT][MT + 4];
  __shared__ float Xs[KT][NT + 4];
  const int b = blockIdx.z;
  const int o0 = blockIdx.y * MT, l0 = blockIdx.x * NT;
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const float *__restrict__ xb = x + (size_t)b * Cin * L;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < Cin; k0 += KT) {
    // W tile: 64 x 16, thread -> (o = t/4, 4 consecutive k)
    {
      const int o = t >> 2, kk = (t & 3) * 4;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + kk + u;
        Ws[kk + u][o] = (o0 + o < Cout && k < Cin) ? __ldg(W + (size_t)(o0 + o) * Cin + k) : 0.f;
      }
    }
    // X tile: 16 x 64, thread -> (k = t/16, 4 consecutive l)
    {
      const int k = t >> 4, ll = (t & 15) * 4;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int l = l0 + ll + u;
        float xv = (k0 + k < Cin && l < L) ? __ldg(xb + (size_t)(k0 + k) * L + l) : 0.f;
        if (pre_scale && k0 + k < Cin) {  // conditional-BN + ReLU applied to the layer INPUT (ONet decoder, fp32 path)
          const size_t pi = (size_t)b * pre_bstride + (k0 + k);
          xv = fmaxf(fmaf(xv, __ldg(pre_scale + pi), __ldg(pre_shift + pi)), 0.f);
        }
        Xs[k][ll + u] = xv;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      float a[4], c[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Ws[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) c[j] = Xs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], c[j], acc[i][j]);
    }
    __syncthreads();
  }

  const int Lout = L / pool;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int o = o0 + ty * 4 + i;
    const bool ov = o < Cout;
    const float sc = ov ? __ldg(scale + o) : 0.f, sh = ov ? __ldg(shift + o) : 0.f;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int l = l0 + tx * 4 + j;
      float r = fmaf(acc[i][j], sc, sh);
      if (relu) r = fmaxf(r, 0.f);
      if (residual && ov && l < L) r += __ldg(residual + ((size_t)b * Cout + o) * L + l);
      v[j] = (l < L) ? r : -INFINITY;
    }
    if (pool == 1) {
      if (ov) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int l = l0 + tx * 4 + j;
          if (l < L) y[((size_t)b * Cout + o) * L + l] = v[j];
        }
      }
    } else {
      float mx = fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3]));
      for (int off = 1; off < pool / 4; off <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const int l = l0 + tx * 4;
      if (ov && (tx % (pool / 4)) == 0 && l < L) y[((size_t)b * Cout + o) * Lout + l / pool] = mx;
    }
  }
}

// out (R^3,3): reference make_3d_grid (external/common.py:157-176): torch.linspace(-0.5,0.5,R) per axis,
// x slowest / z fastest, then * box_size (generator.py:92-95).  torch.linspace (fp32, CUDA and CPU) evaluates
// step = (end-start)/(steps-1) and  v[i] = i < steps/2 ? start + step*i : end - step*(steps-1-i).
__global__ void make_3d_grid_kernel(int R, float box, float *__restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = R * R * R;
  if (e >= total) return;
  const int iz = e % R, iy = (e / R) % R, ix = e / (R * R);
  const float start = -0.5f, end = 0.5f;
  const float step = R > 1 ? __fdiv_rn(__fsub_rn(end, start), (float)(R - 1)) : 0.f;
  const int half = R / 2;
  // the reference builds the lattice on the CPU (generator.py:92-95); torch's CPU linspace at R = 32 rounds the
  // product and the sum separately (no FMA) -- pinned by tests/golden/grid32.npy.
  auto lin = [&](int i) {
    return i < half ? __fadd_rn(start, __fmul_rn(step, (float)i)) : __fsub_rn(end, __fmul_rn(step, (float)(R - 1 - i)));
  };
  out[(size_t)e * 3 + 0] = __fmul_rn(box, lin(ix));
  out[(size_t)e * 3 + 1] = __fmul_rn(box, lin(iy));
  out[(size_t)e * 3 + 2] = __fmul_rn(box, lin(iz));
}

// occupancy bit mask: bit (t % 32) of word (t / 32) of object b is set iff logits[b][t] >= threshold; counts[b] = number
// of occupied points.  First step of the dense-grid -> surface hand-off: callers that only need
// occupancy (voxel IoU, external/common.py:7-35; the `logit >= 0` decision of generator.py:160) copy 1 bit instead of
// 32 per query point back to the host.
__global__ void __launch_bounds__(256)
occupancy_bits_kernel(const float *__restrict__ logits, int T, float threshold, uint32_t *__restrict__ bits,
                      int *__restrict__ counts) {
  const int b = blockIdx.y;
  const int words = (T + 31) / 32;
  const float *lg = logits + (size_t)b * T;
  int local = 0;
  for (int t0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; t0 < words * 32; t0 += gridDim.x * blockDim.x) {
    const int t = t0 + (threadIdx.x & 31);
    const bool occ = t < T && __ldg(lg + t) >= threshold;
    const unsigned m = __ballot_sync(0xffffffffu, occ);
    if ((threadIdx.x & 31) == 0) {
      bits[(size_t)b * words + t0 / 32] = m;
      local += __popc(m);
    }
  }
  if (counts && (threadIdx.x & 31) == 0 && local) atomicAdd(counts + b, local);
}

}  // namespace rfd

using namespace rfd;

extern "C" int rfd_occupancy_bits(const float *logits, int B, int T, float threshold, uint32_t *bits, int *counts,
                                  void *stream) {
  if (B < 0 || T < 0) return RFD_ERR_INVALID_ARGUMENT;
  if (B == 0 || T == 0) return RFD_OK;
  if (!logits || !bits) return RFD_ERR_INVALID_ARGUMENT;
  if (B > 65535) return RFD_ERR_UNSUPPORTED_SIZE;
  cudaStream_t st = as_stream(stream);
  if (counts) RFD_CHECK_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * (size_t)B, st), "occupancy_bits memset");
  int gx = h_ceil_div(T, 256);
  if (gx > 64) gx = 64;
  occupancy_bits_kernel<<<dim3(gx, B), 256, 0, st>>>(logits, T, threshold, bits, counts);
  RFD_CHECK_LAUNCH("occupancy_bits_kernel");
  return RFD_OK;
}

int rfd::launch_pointwise_f32(const float *x, const float *W, const float *scale, const float *shift,
                              const float *residual, const float *pre_scale, const float *pre_shift,
                              long long pre_bstride, int relu, int pool, int B, int Cin, int Cout, int L, float *y,
                              cudaStream_t stream) {
  dim3 grid(h_ceil_div(L, NT), h_ceil_div(Cout, MT), B);
  pointwise_mlp_f32_kernel<<<grid, 256, 0, stream>>>(x, W, scale, shift, residual, pre_scale, pre_shift, pre_bstride,
                                                     relu, pool, Cin, Cout, L, y);
  RFD_CHECK_LAUNCH("pointwise_mlp_f32_kernel");
  return RFD_OK;
}

extern "C" int rfd_pointwise_mlp_f32(const float *x, const float *W, const float *scale, const float *shift,
                                     const float *residual, int relu, int pool, int B, int Cin, int Cout, int L,
                                     float *y, void *stream) {
  if (B < 0 || Cin < 1 || Cout < 1 || L < 0 || pool < 1) return RFD_ERR_INVALID_ARGUMENT;
  if (B == 0 || L == 0) return RFD_OK;
  if (!x || !W || !scale || !shift || !y) return RFD_ERR_INVALID_ARGUMENT;
  if (pool != 1 && !(pool == 4 || pool == 8 || pool == 16 || pool == 32 || pool == 64)) return RFD_ERR_UNSUPPORTED_SIZE;
  if (L % pool != 0 || (pool > 1 && residual)) return RFD_ERR_INVALID_ARGUMENT;
  if (B > 65535 || h_ceil_div(Cout, MT) > 65535) return RFD_ERR_UNSUPPORTED_SIZE;
  return launch_pointwise_f32(x, W, scale, shift, residual, nullptr, nullptr, 0, relu, pool, B, Cin, Cout, L, y,
                             as_stream(stream));
}

extern "C" int rfd_make_3d_grid(int R, float box_size, float *out, void *stream) {
  if (R < 1 || R > 1024 || !out) return RFD_ERR_INVALID_ARGUMENT;
  const int total = R * R * R;
  make_3d_grid_kernel<<<h_ceil_div(total, 256), 256, 0, as_stream(stream)>>>(R, box_size, out);
  RFD_CHECK_LAUNCH("make_3d_grid_kernel");
  return RFD_OK;
}
