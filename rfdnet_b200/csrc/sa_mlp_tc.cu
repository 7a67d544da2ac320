// sa_mlp_tc.cu -- the grouped shared-MLP (+ max over nsample) of a set-abstraction layer on tcgen05 tensor cores.
//
// Reference: build_shared_mlp + F.max_pool2d in PointnetSAModuleVotes.forward
// (external/pointnet2_ops_lib/pointnet2_ops/pointnet2_modules.py:9-19, 237-243): three
// Conv2d(1x1, bias=False) + BatchNorm2d + ReLU over the grouped tensor (B, 3+C, npoint, nsample), then a max over
// nsample -- 10 cuDNN/elementwise launches and three full-size intermediates per layer.
//
// Here (eval mode, BN folded: scale into the bf16 weights, shift kept in fp32) the three GEMMs of a tile of 128
// (point, sample) rows are chained inside one persistent, warp-specialised kernel:
//   grouped tile (fp32, channel-major in HBM) -> bf16 K-major 128B-swizzled A panels in shared memory
//   -> tcgen05.mma (M=128, N=C1) -> TMEM -> +shift, ReLU, bf16 -> A panels -> tcgen05.mma (N=C2) -> ... (N=C3)
//   -> +shift, ReLU, max over the S rows of every group (warp shuffles) -> out (B, C3, npoint) fp32.
// No intermediate activation ever touches HBM; the only HBM traffic is the grouped tensor read (once) and the pooled
// output.  Weights (<= 180 KB bf16 per layer set) stream from L2 through a 4-slot ring of bulk copies like the
// decoder's.  Numerics: bf16 operands, fp32 accumulation (config 3 "bf16 tensor-core path"); the fp32-exact path
// stays in mlp_f32.cu.
#include "common.cuh"
#include "umma.cuh"

namespace rfd {

constexpr int SA_TILE_M = 128;
constexpr int SA_MAX_KP0 = 5;                       // input channels <= 320
constexpr int SA_PANEL = SA_TILE_M * 128;           // 16 KB
constexpr int SA_SLOT = 256 * 128;                  // weight ring slot: up to 256 output rows x 64 k
constexpr int SA_NSLOT = 4;
constexpr int SA_EPI_WARPS = 16;
constexpr int SA_THREADS = 64 + 32 * SA_EPI_WARPS;
constexpr int SA_SM_A = 0;
constexpr int SA_SM_W = SA_SM_A + SA_MAX_KP0 * SA_PANEL;        // 81920
constexpr int SA_SM_SHIFT = SA_SM_W + SA_NSLOT * SA_SLOT;       // 212992
constexpr int SA_SM_BAR = SA_SM_SHIFT + 512 * 4;                // 215040
constexpr int SA_SMEM_BYTES = SA_SM_BAR + 256 + 1024;

struct SaBars {
  uint64_t w_full[SA_NSLOT];
  uint64_t w_empty[SA_NSLOT];
  uint64_t a_ready;
  uint64_t acc_ready;
  uint32_t tmem_base;
};

struct SaParams {
  // gather mode (idx != nullptr): the grouped tile is never materialised -- rows are gathered straight from the
  // point cloud: channel c < 3: (xyz[idx][c] - new_xyz[m][c]) * inv_r ; c >= 3: features[c-3][idx]
  const int *idx;        // (B, M, S) ball-query result, or nullptr
  const float *xyz;      // (B, N, 3)
  const float *new_xyz;  // (B, M, 3)
  const float *feat;     // (B, Ct-3, N) or nullptr
  int N;
  float inv_r;           // 1/radius when normalize_xyz, else 1
  const float *x;        // (B, Ct, L) grouped tensor, L = M*S  (materialised mode)
  const uint8_t *w;      // packed weights: stages (layer, k-panel), each n[layer]*128 bytes
  const float *shift;    // [C1 + C2 + C3]
  float *out;            // (B, C3, M)
  int B, Ct, L, M, S;
  int kp[3];             // k panels per layer
  int n[3];              // output widths C1, C2, C3
  int tiles_per_scene, num_tiles;
};

__global__ void __launch_bounds__(SA_THREADS, 1) sa_mlp_tc_kernel(const SaParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *s_a = smem + SA_SM_A;
  uint8_t *s_w = smem + SA_SM_W;
  float *s_shift = reinterpret_cast<float *>(smem + SA_SM_SHIFT);
  SaBars *bars = reinterpret_cast<SaBars *>(smem + SA_SM_BAR);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
    for (int i = 0; i < SA_NSLOT; ++i) { umma::mbar_init(&bars->w_full[i], 1); umma::mbar_init(&bars->w_empty[i], 1); }
    umma::mbar_init(&bars->a_ready, SA_EPI_WARPS);
    umma::mbar_init(&bars->acc_ready, 1);
    umma::fence_barrier_init();
  }
  if (warp == 1) umma::tmem_alloc(&bars->tmem_base, 512);
  for (int e = tid; e < P.n[0] + P.n[1] + P.n[2]; e += SA_THREADS) s_shift[e] = __ldg(P.shift + e);
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const int tile_lo = (int)(((long long)P.num_tiles * blockIdx.x) / gridDim.x);
  const int tile_hi = (int)(((long long)P.num_tiles * (blockIdx.x + 1)) / gridDim.x);

  if (warp == 0) {
    // ---------------- producer: weight stages in consumption order
    if (lane == 0) {
      uint32_t st = 0, ph = 0;
      for (int tile = tile_lo; tile < tile_hi; ++tile) {
        size_t off = 0;
        for (int l = 0; l < 3; ++l) {
          const uint32_t bytes = (uint32_t)P.n[l] * 128u;
          for (int kp = 0; kp < P.kp[l]; ++kp) {
            umma::mbar_wait(&bars->w_empty[st], ph ^ 1u);
            umma::mbar_arrive_expect_tx(&bars->w_full[st], bytes);
            umma::bulk_g2s(s_w + st * SA_SLOT, P.w + off, bytes, &bars->w_full[st]);
            off += bytes;
            if (++st == SA_NSLOT) { st = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer
    if (lane == 0) {
      const uint32_t a_addr = umma::smem_u32(s_a), w_addr = umma::smem_u32(s_w);
      uint32_t st = 0, ph = 0, step = 0;
      for (int tile = tile_lo; tile < tile_hi; ++tile) {
        for (int l = 0; l < 3; ++l, ++step) {
          const uint32_t idesc = umma::make_idesc_bf16_f32(SA_TILE_M, (uint32_t)P.n[l]);
          umma::mbar_wait(&bars->a_ready, step & 1u);
          for (int kp = 0; kp < P.kp[l]; ++kp) {
            umma::mbar_wait(&bars->w_full[st], ph);
            umma::tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma::mma_bf16_ss(tmem_base + (uint32_t)(l == 0 ? 0 : (l == 1 ? 128 : 256)), umma::make_desc_k_sw128(a_addr + kp * SA_PANEL + k * 32),
                                umma::make_desc_k_sw128(w_addr + st * SA_SLOT + k * 32), idesc, (uint32_t)((kp | k) != 0));
            umma::mma_commit(&bars->w_empty[st]);
            if (++st == SA_NSLOT) { st = 0; ph ^= 1u; }
          }
          umma::mma_commit(&bars->acc_ready);
        }
      }
    }
  } else {
    // ---------------- loader + epilogue warps (16)
    const int et = tid - 64;          // 0..511
    const int q = warp & 3;           // TMEM lane quarter
    const int cq = (warp - 2) >> 2;   // 16-column quarter of every 64-column panel
    const int lr = lane >> 2, lc = lane & 3;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t a_base = umma::smem_u32(s_a);
    const uint32_t pan0 = a_base + (q * 32 + lr) * 128 + lc * 4;
    const uint32_t shift_a = umma::smem_u32(s_shift);
    uint32_t step = 0;
    for (int tile = tile_lo; tile < tile_hi; ++tile) {
      const int b = tile / P.tiles_per_scene;
      const int l0 = (tile - b * P.tiles_per_scene) * SA_TILE_M;
      // ---- A0: grouped tile -> bf16 swizzled panels.  thread = (row r, 8-channel chunk c8); a warp covers 32
      // consecutive rows of one chunk: coalesced 128-B global reads, conflict-free 16-B shared stores.
      {
        const int r = et & 127;
        const bool rv = (l0 + r) < P.L;
        int pk = 0;
        float cx = 0.f, cy = 0.f, cz = 0.f;
        if (P.idx && rv) {
          pk = __ldg(P.idx + (size_t)b * P.L + l0 + r);
          const float *nc = P.new_xyz + ((size_t)b * P.M + (l0 + r) / P.S) * 3;
          cx = __ldg(nc); cy = __ldg(nc + 1); cz = __ldg(nc + 2);
        }
        const float *xb = P.idx ? nullptr : P.x + (size_t)b * P.Ct * P.L;
        const float *pb = P.idx ? P.xyz + ((size_t)b * P.N + pk) * 3 : nullptr;
        const float *fb = (P.idx && P.feat) ? P.feat + (size_t)b * (P.Ct - 3) * P.N + pk : nullptr;
        for (int c8 = et >> 7; c8 < P.kp[0] * 8; c8 += 4) {
          float f[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int c = c8 * 8 + u;
            float v = 0.f;
            if (rv && c < P.Ct) {
              if (!P.idx) {
                v = __ldg(xb + (size_t)c * P.L + l0 + r);
              } else if (c < 3) {
                // same arithmetic as query_and_group_kernel: (p - centre), then * (1/r)
                v = __fsub_rn(__ldg(pb + c), c == 0 ? cx : (c == 1 ? cy : cz));
                if (P.inv_r != 1.0f) v = __fmul_rn(v, P.inv_r);
              } else {
                v = __ldg(fb + (size_t)(c - 3) * P.N);
              }
            }
            f[u] = v;
          }
          const uint32_t dst = a_base + (c8 >> 3) * SA_PANEL + r * 128 + ((((c8 & 7) ^ (r & 7))) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(umma::pack_bf16x2(f[0], f[1])),
                       "r"(umma::pack_bf16x2(f[2], f[3])), "r"(umma::pack_bf16x2(f[4], f[5])),
                       "r"(umma::pack_bf16x2(f[6], f[7]))
                       : "memory");
        }
        umma::fence_proxy_async_smem();
        umma::tc_fence_before();
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(&bars->a_ready);
      }
      int shift_off = 0;
#pragma unroll 1
      for (int l = 0; l < 3; ++l, ++step) {
        umma::mbar_wait(&bars->acc_ready, step & 1u);
        umma::tc_fence_after();
        const int npan = P.n[l] >> 6;
        const uint32_t tmem_l = tmem_base + (uint32_t)(l == 0 ? 0 : (l == 1 ? 128 : 256));
        if (l < 2) {
#pragma unroll 1
          for (int pn = 0; pn < npan; ++pn) {
            const int cb = pn * 64 + cq * 16;
            uint32_t v[2][8];
            umma::tmem_ld_16x256b_x2(tmem_l + lane_base + cb, v[0]);
            umma::tmem_ld_16x256b_x2(tmem_l + lane_base + (16u << 16) + cb, v[1]);
            const float2 s0 = umma::lds_f2(shift_a + (shift_off + cb + 2 * lc) * 4);
            const float2 s1 = umma::lds_f2(shift_a + (shift_off + cb + 8 + 2 * lc) * 4);
            umma::tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const float2 sh = i ? s1 : s0;
              const uint32_t sw = (uint32_t)(((cq * 2 + i) ^ lr) << 4);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float x0 = __uint_as_float(v[j >> 1][4 * i + 2 * (j & 1)]) + sh.x;
                const float x1 = __uint_as_float(v[j >> 1][4 * i + 2 * (j & 1) + 1]) + sh.y;
                umma::sts_u32(pan0 + pn * SA_PANEL + j * 1024 + sw, umma::pack_relu_bf16x2(x0, x1));
              }
            }
          }
          umma::fence_proxy_async_smem();
          umma::tc_fence_before();
          __syncwarp();
          if (lane == 0) umma::mbar_arrive(&bars->a_ready);
        } else {
          // ---- final layer: ReLU, max over the S rows of every group, store (B, C3, M)
          const int gpt = SA_TILE_M / P.S;  // groups per tile
          const int m0 = (l0 / P.S);
          float *ob = P.out + (size_t)b * P.n[2] * P.M;
#pragma unroll 1
          for (int pn = 0; pn < npan; ++pn) {
            const int cb = pn * 64 + cq * 16;
            uint32_t v[2][8];
            umma::tmem_ld_16x256b_x2(tmem_l + lane_base + cb, v[0]);
            umma::tmem_ld_16x256b_x2(tmem_l + lane_base + (16u << 16) + cb, v[1]);
            const float2 s0 = umma::lds_f2(shift_a + (shift_off + cb + 2 * lc) * 4);
            const float2 s1 = umma::lds_f2(shift_a + (shift_off + cb + 8 + 2 * lc) * 4);
            umma::tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const float2 sh = i ? s1 : s0;
              // rows of this thread: 32q + lr + 8j.  S=16: {j=0,1} and {j=2,3} are two groups; S>=32: one.
              float g0[2], g1[2];
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const float shv = u ? sh.y : sh.x;
                const float r0 = fmaxf(__uint_as_float(v[0][4 * i + u]) + shv, 0.f);
                const float r1 = fmaxf(__uint_as_float(v[0][4 * i + 2 + u]) + shv, 0.f);
                const float r2 = fmaxf(__uint_as_float(v[1][4 * i + u]) + shv, 0.f);
                const float r3 = fmaxf(__uint_as_float(v[1][4 * i + 2 + u]) + shv, 0.f);
                g0[u] = fmaxf(r0, r1);
                g1[u] = fmaxf(r2, r3);
                if (P.S >= 32) { g0[u] = fmaxf(g0[u], g1[u]); g1[u] = g0[u]; }
#pragma unroll
                for (int off = 4; off <= 16; off <<= 1) {
                  g0[u] = fmaxf(g0[u], __shfl_xor_sync(0xffffffffu, g0[u], off));
                  g1[u] = fmaxf(g1[u], __shfl_xor_sync(0xffffffffu, g1[u], off));
                }
              }
              if (lr == 0) {
                const int col = cb + 8 * i + 2 * lc;
                if (P.S == 16) {
                  const int ga = m0 + 2 * q, gb = ga + 1;
                  if (ga < P.M) { ob[(size_t)col * P.M + ga] = g0[0]; ob[(size_t)(col + 1) * P.M + ga] = g0[1]; }
                  if (gb < P.M) { ob[(size_t)col * P.M + gb] = g1[0]; ob[(size_t)(col + 1) * P.M + gb] = g1[1]; }
                } else if (P.S == 32) {
                  const int ga = m0 + q;
                  if (ga < P.M) { ob[(size_t)col * P.M + ga] = g0[0]; ob[(size_t)(col + 1) * P.M + ga] = g0[1]; }
                } else {  // S == 64 (two lane quarters per group) or 128: values are >= 0, so integer max == float max
                  const int ga = m0 + (q * 32) / P.S;
                  if (ga < P.M) {
                    atomicMax(reinterpret_cast<int *>(ob + (size_t)col * P.M + ga), __float_as_int(g0[0]));
                    atomicMax(reinterpret_cast<int *>(ob + (size_t)(col + 1) * P.M + ga), __float_as_int(g0[1]));
                  }
                }
              }
            }
          }
          (void)gpt;
          umma::tc_fence_before();
        }
        shift_off += P.n[l];
      }
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    umma::tc_fence_after();
    umma::tmem_dealloc(tmem_base, 512);
  }
}

// pack one layer: W (N, K) f32 row-major with the BN scale folded in -> bf16 stages [kp][N rows][64 k] swizzled
__global__ void sa_pack_kernel(const float *__restrict__ W, const float *__restrict__ scale, int N, int K, int kpn,
                               uint8_t *__restrict__ dst) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;  // one thread per 16-byte chunk
  const int total = kpn * N * 8;
  if (e >= total) return;
  const int kp = e / (N * 8), rem = e % (N * 8);
  const int n = rem / 8, cin = rem % 8;
  const float sc = __ldg(scale + n);
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k0 = kp * 64 + cin * 8 + 2 * i;
    const float a = k0 < K ? __ldg(W + (size_t)n * K + k0) * sc : 0.f;
    const float b = k0 + 1 < K ? __ldg(W + (size_t)n * K + k0 + 1) * sc : 0.f;
    w[i] = umma::pack_bf16x2(a, b);
  }
  *reinterpret_cast<uint4 *>(dst + (size_t)kp * N * 128 + n * 128 + ((cin ^ (n & 7)) << 4)) =
      make_uint4(w[0], w[1], w[2], w[3]);
}

}  // namespace rfd

using namespace rfd;

static int sa_kp0(int Ct) { return (Ct + 63) / 64; }

extern "C" size_t rfd_sa_mlp_tc_packed_bytes(int Ct, int C1, int C2, int C3) {
  if (Ct < 1 || sa_kp0(Ct) > SA_MAX_KP0 || C1 % 64 || C2 % 64 || C3 % 64 || C1 < 64 || C2 < 64 || C3 < 64 ||
      C1 > 128 || C2 > 128 || C3 > 256)
    return 0;
  return (size_t)sa_kp0(Ct) * C1 * 128 + (size_t)(C1 / 64) * C2 * 128 + (size_t)(C2 / 64) * C3 * 128;
}

extern "C" int rfd_sa_mlp_tc_pack(const float *W1, const float *scale1, const float *W2, const float *scale2,
                                  const float *W3, const float *scale3, int Ct, int C1, int C2, int C3, void *packed,
                                  void *stream) {
  if (!W1 || !W2 || !W3 || !scale1 || !scale2 || !scale3 || !packed) return RFD_ERR_INVALID_ARGUMENT;
  if (rfd_sa_mlp_tc_packed_bytes(Ct, C1, C2, C3) == 0) return RFD_ERR_UNSUPPORTED_SIZE;
  uint8_t *dst = reinterpret_cast<uint8_t *>(packed);
  const float *Ws[3] = {W1, W2, W3};
  const float *Ss[3] = {scale1, scale2, scale3};
  const int Ks[3] = {Ct, C1, C2}, Ns[3] = {C1, C2, C3};
  for (int l = 0; l < 3; ++l) {
    const int kpn = (Ks[l] + 63) / 64;
    const int total = kpn * Ns[l] * 8;
    sa_pack_kernel<<<h_ceil_div(total, 256), 256, 0, as_stream(stream)>>>(Ws[l], Ss[l], Ns[l], Ks[l], kpn, dst);
    RFD_CHECK_LAUNCH("sa_pack_kernel");
    dst += (size_t)kpn * Ns[l] * 128;
  }
  return RFD_OK;
}

static int sa_launch(SaParams &P, int B, int Ct, int M, int S, const void *packed, const float *shift, int C1, int C2,
                     int C3, float *out, void *stream) {
  if (rfd_sa_mlp_tc_packed_bytes(Ct, C1, C2, C3) == 0) return RFD_ERR_UNSUPPORTED_SIZE;
  if (!(S == 16 || S == 32 || S == 64 || S == 128)) return RFD_ERR_UNSUPPORTED_SIZE;
  P.w = reinterpret_cast<const uint8_t *>(packed); P.shift = shift; P.out = out;
  P.B = B; P.Ct = Ct; P.M = M; P.S = S; P.L = M * S;
  P.kp[0] = sa_kp0(Ct); P.kp[1] = C1 / 64; P.kp[2] = C2 / 64;
  P.n[0] = C1; P.n[1] = C2; P.n[2] = C3;
  P.tiles_per_scene = (P.L + SA_TILE_M - 1) / SA_TILE_M;
  const long long nt = (long long)P.tiles_per_scene * B;
  if (nt > 0x7fffffffLL) return RFD_ERR_UNSUPPORTED_SIZE;
  P.num_tiles = (int)nt;
  cudaStream_t st = as_stream(stream);
  if (S > 32) RFD_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)B * C3 * M, st), "sa_mlp_tc memset");
  int dev = 0, sms = 148;
  RFD_CHECK_CUDA(cudaGetDevice(&dev), "sa_mlp_tc getdevice");
  RFD_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "sa_mlp_tc sms");
  RFD_CHECK_CUDA(cudaFuncSetAttribute(sa_mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SA_SMEM_BYTES),
                 "sa_mlp_tc attr");
  const int grid = (int)(nt < sms ? nt : sms);
  sa_mlp_tc_kernel<<<grid, SA_THREADS, SA_SMEM_BYTES, st>>>(P);
  RFD_CHECK_LAUNCH("sa_mlp_tc_kernel");
  return RFD_OK;
}

extern "C" int rfd_sa_mlp_tc(const float *x, int B, int Ct, int M, int S, const void *packed, const float *shift,
                             int C1, int C2, int C3, float *out, void *stream) {
  if (B < 0 || M < 0 || S < 1) return RFD_ERR_INVALID_ARGUMENT;
  if (B == 0 || M == 0) return RFD_OK;
  if (!x || !packed || !shift || !out) return RFD_ERR_INVALID_ARGUMENT;
  SaParams P = {};
  P.x = x;
  return sa_launch(P, B, Ct, M, S, packed, shift, C1, C2, C3, out, stream);
}

extern "C" int rfd_sa_gather_mlp_tc(const float *xyz, const float *new_xyz, const float *features, const int *idx, int B,
                                    int N, int M, int S, int C, float radius, int normalize_xyz, const void *packed,
                                    const float *shift, int C1, int C2, int C3, float *out, void *stream) {
  if (B < 0 || M < 0 || S < 1 || N < 1 || C < 0) return RFD_ERR_INVALID_ARGUMENT;
  if (B == 0 || M == 0) return RFD_OK;
  if (!xyz || !new_xyz || !idx || (C > 0 && !features) || !packed || !shift || !out) return RFD_ERR_INVALID_ARGUMENT;
  SaParams P = {};
  P.idx = idx; P.xyz = xyz; P.new_xyz = new_xyz; P.feat = C > 0 ? features : nullptr; P.N = N;
  P.inv_r = normalize_xyz ? 1.0f / radius : 1.0f;
  return sa_launch(P, B, 3 + C, M, S, packed, shift, C1, C2, C3, out, stream);
}
