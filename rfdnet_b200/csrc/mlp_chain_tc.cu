// mlp_chain_tc.cu -- every dense pointwise-MLP of the detection pass on tcgen05 tensor cores, one kernel per module.
//
// Reference (eval mode; external/pointnet2_ops_lib/pointnet2_ops/pointnet2_modules.py, models/iscnet/modules/):
//   PointnetSAModuleVotes  pointnet2_modules.py:9-19,237-243   3 x {Conv2d 1x1 (no bias) + BatchNorm2d + ReLU} over the
//                                                              grouped tensor (B,3+C,npoint,nsample), then max over nsample
//   PointnetFPModule       pointnet2_modules.py:395-405        2 x {Conv2d 1x1 + BN + ReLU} on (B,512,n,1)
//   VotingModule           vote_module.py:34-61                conv1d+BN+ReLU x2, conv1d 256 -> 259
//   ProposalModule head    proposal_module.py:85-124           conv1d+BN+ReLU x2, conv1d 128 -> 69
// i.e. 10 cuDNN/elementwise launches and full-size intermediates per module in the reference.
//
// Here a module is ONE persistent, warp-specialised kernel.  Per tile of 128 rows (a row = one (point, sample) slot of
// an SA layer, or one point of an FP / head layer):
//   loader     16 warps build the layer-0 A operand in shared memory (K-major, 128B-swizzled 16-bit panels), either
//                * gather mode: rows are gathered through the ball-query `idx` straight from the POINT-MAJOR feature
//                  tensor (B,N,C) -- a slot is one contiguous row, read with 16-byte loads -- the (B,3+C,M,S) grouped
//                  tensor of the reference is never materialised; or
//                * dense mode: a channel-major (B,K,L) tensor, coalesced along L;
//   MMA        one thread issues tcgen05.mma (M=128, N=layer width, K=16) per 64-wide K panel; accumulators in TMEM
//              (columns 0 / 256 alternate between layers);
//   epilogue   the same 16 warps read the accumulator (tcgen05.ld 16x256b), apply the folded BatchNorm affine
//              y = scale*acc + shift (fp32) and ReLU, and write the next layer's A panels; the last layer's epilogue
//              max-pools over the nsample rows of each group (warp shuffles) or writes the rows as they are, in
//              channel-major (B,C,L) and/or point-major (B,L,C) layout;
//   weights    stream from L2 through a ring of 32-KB bulk (TMA) copies issued by a producer warp.
// The three relative-xyz input channels of an SA layer never enter the tensor-core GEMM: their contribution
// (3 FMAs per output) is added in fp32 by the layer-0 epilogue from a per-row (dx,dy,dz) table, so K is exactly C.
//
// Precision modes (same as the ONet decoder): bf16 / fp16 single-MMA, and split-fp16 "x3"
// (a_hi.w_hi + a_lo.w_hi + a_hi.w_lo, ~22 significant bits, fp32 accumulate), which holds BASELINE config 2's
// fp32 / 1e-4 parity on tensor cores.  No intermediate activation touches HBM in any mode.
#include "common.cuh"
#include "umma.cuh"

namespace rfd {

constexpr int CH_TILE_M = 128;
constexpr int CH_PANEL = CH_TILE_M * 128;   // 16 KB: 128 rows x 64 16-bit elements
constexpr int CH_APAN = 4;                  // resident K panels of the A operand (x3: 4 hi + 4 lo)
constexpr int CH_SLOT = 256 * 128;          // weight ring slot: up to 256 output rows x 64 k
constexpr int CH_EPI_WARPS = 16;
constexpr int CH_THREADS = 64 + 32 * CH_EPI_WARPS;
constexpr int CH_MAX_STEPS = 4;
constexpr int CH_MODE_BF16 = 1, CH_MODE_F16 = 2, CH_MODE_F16X3 = 3;
constexpr int CH_TAB = 1024;                // floats per scale / shift table
// shared memory map (1024-B aligned base)
constexpr int CH_SM_A = 0;                                   // 128 KB (single modes use the first 64 KB)
constexpr int CH_SM_W = 8 * CH_PANEL;                        // ring: x3 2 slots, single 2 slots
constexpr int CH_NSLOT = 2;
constexpr int CH_SM_SCALE = CH_SM_W + CH_NSLOT * CH_SLOT;    // 196608
constexpr int CH_SM_SHIFT = CH_SM_SCALE + CH_TAB * 4;
constexpr int CH_SM_WXYZ = CH_SM_SHIFT + CH_TAB * 4;         // [256][4] fp32: layer-0 weights of the xyz channels
constexpr int CH_SM_REL = CH_SM_WXYZ + 256 * 16;             // [128][4] fp32: relative xyz of the tile's rows
constexpr int CH_SM_BAR = CH_SM_REL + CH_TILE_M * 16;
constexpr int CH_SMEM_BYTES = CH_SM_BAR + 256 + 1024;
static_assert(CH_SMEM_BYTES <= 232448, "exceeds the 227 KB of dynamic shared memory per CTA");

struct ChainBars {
  uint64_t w_full[CH_NSLOT];
  uint64_t w_empty[CH_NSLOT];
  uint64_t a_ready;      // 16 epilogue/loader warps: an A operand (round of layer 0, or a layer's output) is in place
  uint64_t a_free;       // MMA: the resident A panels of a layer-0 round have been consumed
  uint64_t a_ready_pp[2];  // ping-pong rounds of a streamed layer-0 operand: one barrier per panel pair, so the loader (which
                           // may run a full round ahead) can never advance a barrier two phases past the MMA thread's wait
  uint64_t acc_ready[2]; // MMA: accumulator of a step complete (ping-pong by step parity)
  uint32_t tmem_base;
};

struct ChainStep {
  int kp;         // K panels (64 wide) of this step's A operand
  int n;          // MMA N (multiple of 16, <= 256); intermediates: multiple of 64
  int relu;
  int is_out;     // 1: final epilogue (pool / dense write); 0: writes the next A operand
  int reuse_a;    // 1: same A operand as the previous step (second column block of a wide output layer)
  int out_ch0;    // first output channel written by this step
  int out_valid;  // valid output channels of this step
  int tab_off;    // offset into the scale / shift tables
};

struct ChainParams {
  // gather mode (idx != nullptr)
  const int *idx;         // (B, M, S)
  const float *xyz;       // (B, N, 3)
  const float *new_xyz;   // (B, M, 3)
  const float *feat_pm;   // (B, N, K0) point-major, or nullptr when K0 == 0
  int N;
  float inv_r;            // 1/radius when normalize_xyz
  int normalize;
  // dense mode
  const float *x;         // (B, K0, L) channel-major
  int K0;                 // tensor-core input channels of step 0
  int has_xyz;            // step 0 adds the fp32 xyz term
  const uint8_t *w;       // packed weight stages in consumption order
  const float *w_xyz;     // [n0][4]
  const float *scale, *shift;  // concatenated per step
  int tab_floats;
  float *out_cm;          // (B, out_C, L / pool) or nullptr
  float *out_pm;          // (B, L / pool, out_C) or nullptr
  int out_C;
  int B, L, M, S, pool;   // pool: rows per max-pool group (1 = none)
  int nsteps;
  ChainStep st[CH_MAX_STEPS];
  int tiles_per_scene, num_tiles;
  // dense-mode extras (rfd_mlp_chain_ex)
  int relu_in;            // ReLU applied to the layer-0 input while it is loaded
  const float *gbias;     // (B, L / gbias_rows, n0) per-group pre-activation bias of layer 0: y = scale * (acc + gbias) + shift
  int gbias_rows;         // rows per group (multiple of the 128-row tile)
  float *out_pool;        // (B, out_C, L / pool_rows): max over groups of rows of the (unpooled) output, caller-initialised
  int pool_rows;          //   to -inf; merged with a sign-aware atomic max (values of either sign)
  // row-major ("point-major") dense mode (rfd_mlp_chain_rows): x_pm (B, L, ldi) with the K0 operand channels in columns
  // [0, K0) of every row -- a tile reads 128 contiguous rows (sequential HBM access; the channel-major form reads 512 B per
  // channel, one DRAM page each) -- out_pm rows have stride ldo and start at column out_col0; pool_pm: out_pool is (B, G, out_C)
  const float *x_pm;
  int ldi, ldo, out_col0, pool_pm;
};

// atomic max on a float of either sign (address initialised to -inf or any float)
__device__ __forceinline__ void atomic_max_float(float *addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}

template <int MODE>
__device__ __forceinline__ void chain_store_act(uint32_t addr, float v0, float v1, bool relu) {
  if (relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
  if (MODE == CH_MODE_BF16) {
    umma::sts_u32(addr, umma::pack_bf16x2(v0, v1));
  } else {
    const uint32_t h = umma::pack_f16x2(v0, v1);
    umma::sts_u32(addr, h);
    if (MODE == CH_MODE_F16X3) {
      const float2 hf = umma::unpack_f16x2(h);
      umma::sts_u32(addr + CH_APAN * CH_PANEL, umma::pack_f16x2(v0 - hf.x, v1 - hf.y));
    }
  }
}

// 8 fp32 values -> one 16-byte chunk of the hi panel (and of the lo panel in x3 mode)
template <int MODE>
__device__ __forceinline__ void chain_store_chunk(uint32_t dst, const float (&f)[8]) {
  uint32_t h[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    h[i] = MODE == CH_MODE_BF16 ? umma::pack_bf16x2(f[2 * i], f[2 * i + 1]) : umma::pack_f16x2(f[2 * i], f[2 * i + 1]);
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
  if (MODE == CH_MODE_F16X3) {
    uint32_t l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 hf = umma::unpack_f16x2(h[i]);
      l[i] = umma::pack_f16x2(f[2 * i] - hf.x, f[2 * i + 1] - hf.y);
    }
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + CH_APAN * CH_PANEL), "r"(l[0]), "r"(l[1]), "r"(l[2]),
                 "r"(l[3])
                 : "memory");
  }
}

template <int MODE>
__global__ void __launch_bounds__(CH_THREADS, 1) mlp_chain_tc_kernel(const ChainParams P) {
  constexpr bool X3 = MODE == CH_MODE_F16X3;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *s_a = smem + CH_SM_A;
  uint8_t *s_w = smem + CH_SM_W;
  float *s_scale = reinterpret_cast<float *>(smem + CH_SM_SCALE);
  float *s_shift = reinterpret_cast<float *>(smem + CH_SM_SHIFT);
  float4 *s_wxyz = reinterpret_cast<float4 *>(smem + CH_SM_WXYZ);
  float4 *s_rel = reinterpret_cast<float4 *>(smem + CH_SM_REL);
  ChainBars *bars = reinterpret_cast<ChainBars *>(smem + CH_SM_BAR);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
    for (int i = 0; i < CH_NSLOT; ++i) { umma::mbar_init(&bars->w_full[i], 1); umma::mbar_init(&bars->w_empty[i], 1); }
    umma::mbar_init(&bars->a_ready, CH_EPI_WARPS);
    umma::mbar_init(&bars->a_free, 1);
    umma::mbar_init(&bars->a_ready_pp[0], CH_EPI_WARPS);
    umma::mbar_init(&bars->a_ready_pp[1], CH_EPI_WARPS);
    umma::mbar_init(&bars->acc_ready[0], 1);
    umma::mbar_init(&bars->acc_ready[1], 1);
    umma::fence_barrier_init();
  }
  if (warp == 1) umma::tmem_alloc(&bars->tmem_base, 512);
  for (int e = tid; e < P.tab_floats; e += CH_THREADS) { s_scale[e] = __ldg(P.scale + e); s_shift[e] = __ldg(P.shift + e); }
  if (P.has_xyz)
    for (int e = tid; e < P.st[0].n; e += CH_THREADS) s_wxyz[e] = __ldg(reinterpret_cast<const float4 *>(P.w_xyz) + e);
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const int tile_lo = (int)(((long long)P.num_tiles * blockIdx.x) / gridDim.x);
  const int tile_hi = (int)(((long long)P.num_tiles * (blockIdx.x + 1)) / gridDim.x);
  const int kp0 = P.st[0].kp;
  // layer-0 operand wider than the 4 resident panels: it streams through the panels in rounds.  Rounds are then 2 panels
  // wide and alternate between panels {0,1} and {2,3} (hi; lo 4 panels above), so the loads + conversion of round r+1
  // overlap the MMAs of round r (the first version used 4-panel rounds back to back: load, multiply, load, ...).
  const bool pingpong = kp0 > CH_APAN;
  const int RP = pingpong ? 2 : CH_APAN;
  const int rounds0 = kp0 == 0 ? 0 : (kp0 + RP - 1) / RP;

  if (warp == 0) {
    // ---------------- producer: weight stages in consumption order: step, k panel, hi [, lo]
    if (lane == 0) {
      uint32_t st = 0, ph = 0;
      for (int tile = tile_lo; tile < tile_hi; ++tile) {
        size_t off = 0;
        for (int s = 0; s < P.nsteps; ++s) {
          const uint32_t bytes = (uint32_t)P.st[s].n * 128u;
          const int nst = P.st[s].kp * (X3 ? 2 : 1);
          for (int i = 0; i < nst; ++i) {
            umma::mbar_wait(&bars->w_empty[st], ph ^ 1u);
            umma::mbar_arrive_expect_tx(&bars->w_full[st], bytes);
            umma::bulk_g2s(s_w + st * CH_SLOT, P.w + off, bytes, &bars->w_full[st]);
            off += bytes;
            if (++st == CH_NSLOT) { st = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer
    if (lane == 0) {
      const uint32_t a_addr = umma::smem_u32(s_a), w_addr = umma::smem_u32(s_w);
      uint32_t st = 0, ph = 0, a_cnt = 0, pp_cnt[2] = {0u, 0u};
      for (int tile = tile_lo; tile < tile_hi; ++tile) {
        for (int s = 0; s < P.nsteps; ++s) {
          const ChainStep S = P.st[s];
          const uint32_t idesc = MODE == CH_MODE_BF16 ? umma::make_idesc_bf16_f32(CH_TILE_M, (uint32_t)S.n)
                                                      : umma::make_idesc_f16_f32(CH_TILE_M, (uint32_t)S.n);
          const uint32_t d_tmem = tmem_base + (uint32_t)((s & 1) * 256);
          const int rounds = s == 0 ? rounds0 : 1;
          uint32_t first = 1;
          for (int r = 0; r < rounds; ++r) {
            if (s == 0 && pingpong) { umma::mbar_wait(&bars->a_ready_pp[r & 1], pp_cnt[r & 1] & 1u); ++pp_cnt[r & 1]; }
            else if (!S.reuse_a) { umma::mbar_wait(&bars->a_ready, a_cnt & 1u); ++a_cnt; }
            umma::tc_fence_after();
            const int kpn = s == 0 ? min(RP, kp0 - r * RP) : S.kp;
            const int pb = (s == 0 && pingpong) ? (r & 1) * 2 : 0;
            for (int kp = 0; kp < kpn; ++kp) {
              const uint32_t a_hi = a_addr + (pb + kp) * CH_PANEL, a_lo = a_hi + CH_APAN * CH_PANEL;
              umma::mbar_wait(&bars->w_full[st], ph);
              umma::tc_fence_after();
              const uint32_t w_hi = w_addr + st * CH_SLOT;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma::mma_f16_ss(d_tmem, umma::make_desc_k_sw128(a_hi + k * 32), umma::make_desc_k_sw128(w_hi + k * 32), idesc,
                                 first ? 0u : 1u);
                first = 0;
                if (X3)
                  umma::mma_f16_ss(d_tmem, umma::make_desc_k_sw128(a_lo + k * 32), umma::make_desc_k_sw128(w_hi + k * 32),
                                   idesc, 1u);
              }
              umma::mma_commit(&bars->w_empty[st]);
              if (++st == CH_NSLOT) { st = 0; ph ^= 1u; }
              if (X3) {
                umma::mbar_wait(&bars->w_full[st], ph);
                umma::tc_fence_after();
                const uint32_t w_lo = w_addr + st * CH_SLOT;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma::mma_f16_ss(d_tmem, umma::make_desc_k_sw128(a_hi + k * 32), umma::make_desc_k_sw128(w_lo + k * 32),
                                   idesc, 1u);
                umma::mma_commit(&bars->w_empty[st]);
                if (++st == CH_NSLOT) { st = 0; ph ^= 1u; }
              }
            }
            if (s == 0 && r + (pingpong ? 2 : 1) < rounds) umma::mma_commit(&bars->a_free);  // a later round reuses these panels
          }
          umma::mma_commit(&bars->acc_ready[s & 1]);
        }
      }
    }
  } else {
    // ---------------- loader + epilogue warps (16)
    const int et = tid - 64;          // 0..511
    const int we = warp - 2;          // 0..15
    const int q = warp & 3;           // TMEM lane quarter (= warp id % 4)
    const int cq = we >> 2;           // 16-column quarter of every 64-column panel
    const int lr = lane >> 2, lc = lane & 3;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t a_base = umma::smem_u32(s_a);
    const uint32_t pan0 = a_base + (q * 32 + lr) * 128 + lc * 4;
    const uint32_t scale_a = umma::smem_u32(s_scale), shift_a = umma::smem_u32(s_shift);
    uint32_t acc_uses[2] = {0u, 0u}, free_cnt = 0;
    for (int tile = tile_lo; tile < tile_hi; ++tile) {
      const int b = tile / P.tiles_per_scene;
      const int l0 = (tile - b * P.tiles_per_scene) * CH_TILE_M;
      // ---- relative xyz of the tile's rows (gather mode): thread et < 128 owns row et
      int pk_row = 0;
      if (P.idx && et < CH_TILE_M) {
        float4 rel = make_float4(0.f, 0.f, 0.f, 0.f);
        if (l0 + et < P.L) {
          pk_row = __ldg(P.idx + (size_t)b * P.L + l0 + et);
          if (P.has_xyz) {
            const float *nc = P.new_xyz + ((size_t)b * P.M + (l0 + et) / P.S) * 3;
            const float *pp = P.xyz + ((size_t)b * P.N + pk_row) * 3;
            // same arithmetic as query_and_group_kernel: (p - centre), then * (1/r)
            rel.x = __fsub_rn(__ldg(pp), __ldg(nc));
            rel.y = __fsub_rn(__ldg(pp + 1), __ldg(nc + 1));
            rel.z = __fsub_rn(__ldg(pp + 2), __ldg(nc + 2));
            if (P.normalize) { rel.x = __fmul_rn(rel.x, P.inv_r); rel.y = __fmul_rn(rel.y, P.inv_r); rel.z = __fmul_rn(rel.z, P.inv_r); }
          }
        }
        s_rel[et] = rel;
      }
      // ---- layer-0 A operand, CH_APAN panels per round
      for (int r = 0; r < rounds0; ++r) {
        if (r >= (pingpong ? 2 : 1)) { umma::mbar_wait(&bars->a_free, free_cnt & 1u); ++free_cnt; }
        const int kpn = min(RP, kp0 - r * RP);
        const int pb = pingpong ? (r & 1) * 2 : 0;
        if (P.idx) {
          // gather: lane = (row sub-index, 8-channel chunk); a warp instruction covers 4 rows x 64 channels (256 B each)
          const int rsub = lane >> 3, ch = lane & 7;
          const float *fb = P.feat_pm + (size_t)b * P.N * P.K0;
#pragma unroll 1
          for (int it = 0; it < 2; ++it) {
            const int row = 4 * (we + 16 * it) + rsub;
            const bool rv = (l0 + row) < P.L;
            const int pk = rv ? __ldg(P.idx + (size_t)b * P.L + l0 + row) : 0;
            const float *src = fb + (size_t)pk * P.K0;
            for (int kp = 0; kp < kpn; ++kp) {
              const int c = (r * RP + kp) * 64 + ch * 8;
              float f[8];
              if (rv && c + 8 <= P.K0 && (P.K0 & 3) == 0) {
                const float4 a = __ldg(reinterpret_cast<const float4 *>(src + c));
                const float4 d = __ldg(reinterpret_cast<const float4 *>(src + c + 4));
                f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = d.x; f[5] = d.y; f[6] = d.z; f[7] = d.w;
              } else {
#pragma unroll
                for (int u = 0; u < 8; ++u) f[u] = (rv && c + u < P.K0) ? __ldg(src + c + u) : 0.f;
              }
              chain_store_chunk<MODE>(a_base + (pb + kp) * CH_PANEL + row * 128 + ((ch ^ (row & 7)) << 4), f);
            }
          }
        } else if (P.x_pm) {
          // dense row-major: lane = (row sub-index, 8-channel chunk); a warp instruction covers 4 rows x 64 channels (256 B each)
          const int rsub = lane >> 3, ch = lane & 7;
          const float *fb = P.x_pm + (size_t)b * P.L * P.ldi;
#pragma unroll 1
          for (int it = 0; it < 2; ++it) {
            const int row = 4 * (we + 16 * it) + rsub;
            const bool rv = (l0 + row) < P.L;
            const float *src = fb + (size_t)(l0 + row) * P.ldi;
            for (int kp = 0; kp < kpn; ++kp) {
              const int c = (r * RP + kp) * 64 + ch * 8;
              float f[8];
              if (rv && c + 8 <= P.K0) {   // ldi and the base are multiples of 4 floats (checked on the host)
                const float4 a = __ldg(reinterpret_cast<const float4 *>(src + c));
                const float4 d = __ldg(reinterpret_cast<const float4 *>(src + c + 4));
                f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = d.x; f[5] = d.y; f[6] = d.z; f[7] = d.w;
              } else {
#pragma unroll
                for (int u = 0; u < 8; ++u) f[u] = (rv && c + u < P.K0) ? __ldg(src + c + u) : 0.f;
              }
              if (P.relu_in) {
#pragma unroll
                for (int u = 0; u < 8; ++u) f[u] = fmaxf(f[u], 0.f);
              }
              chain_store_chunk<MODE>(a_base + (pb + kp) * CH_PANEL + row * 128 + ((ch ^ (row & 7)) << 4), f);
            }
          }
        } else {
          // dense channel-major: thread = (row, 8-channel chunk); a warp covers 32 consecutive rows of one chunk
          const int row = et & 127;
          const bool rv = (l0 + row) < P.L;
          const float *xb = P.x + (size_t)b * P.K0 * P.L + l0 + row;
          for (int c8 = et >> 7; c8 < kpn * 8; c8 += 4) {
            float f[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int c = r * RP * 64 + c8 * 8 + u;
              f[u] = (rv && c < P.K0) ? __ldg(xb + (size_t)c * P.L) : 0.f;
              if (P.relu_in) f[u] = fmaxf(f[u], 0.f);
            }
            chain_store_chunk<MODE>(a_base + (pb + (c8 >> 3)) * CH_PANEL + row * 128 + (((c8 & 7) ^ (row & 7)) << 4), f);
          }
        }
        umma::fence_proxy_async_smem();
        umma::tc_fence_before();
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(pingpong ? &bars->a_ready_pp[r & 1] : &bars->a_ready);
      }
      if (P.has_xyz) asm volatile("bar.sync 1, %0;" ::"n"(32 * CH_EPI_WARPS) : "memory");  // s_rel visible to all epilogue warps
      (void)pk_row;
#pragma unroll 1
      for (int s = 0; s < P.nsteps; ++s) {
        const ChainStep S = P.st[s];
        umma::mbar_wait(&bars->acc_ready[s & 1], acc_uses[s & 1] & 1u);
        ++acc_uses[s & 1];
        umma::tc_fence_after();
        const int npan = (S.n + 63) >> 6;
        const uint32_t tmem_s = tmem_base + (uint32_t)((s & 1) * 256);
        const bool xyz_term = s == 0 && P.has_xyz;
        const bool no_acc = s == 0 && kp0 == 0;  // xyz-only layer: nothing was accumulated in TMEM
        // per-group bias of layer 0 (a tile never straddles groups: gbias_rows is a multiple of the tile height)
        const float *gbp = (s == 0 && P.gbias) ? P.gbias + ((size_t)b * (P.L / P.gbias_rows) + l0 / P.gbias_rows) * S.n : nullptr;
        float4 rel[4];
        if (xyz_term) {
#pragma unroll
          for (int j = 0; j < 4; ++j) rel[j] = s_rel[q * 32 + lr + 8 * j];
        }
        const int gpool = P.pool;
        const int m0 = l0 / gpool;          // first output row (group) of the tile
        const int Lo = P.L / gpool;         // output rows per scene
#pragma unroll 1
        for (int pn = 0; pn < npan; ++pn) {
          const int cb = pn * 64 + cq * 16;
          if (cb >= S.n) continue;  // warp-uniform
          uint32_t v[2][8];
          umma::tmem_ld_16x256b_x2(tmem_s + lane_base + cb, v[0]);
          umma::tmem_ld_16x256b_x2(tmem_s + lane_base + (16u << 16) + cb, v[1]);
          const int t0 = S.tab_off + cb + 2 * lc;
          const float2 a0 = umma::lds_f2(scale_a + t0 * 4), a1 = umma::lds_f2(scale_a + (t0 + 8) * 4);
          const float2 s0 = umma::lds_f2(shift_a + t0 * 4), s1 = umma::lds_f2(shift_a + (t0 + 8) * 4);
          float2 gb0 = make_float2(0.f, 0.f), gb1 = gb0;
          if (gbp) {
            gb0 = __ldg(reinterpret_cast<const float2 *>(gbp + cb + 2 * lc));
            gb1 = __ldg(reinterpret_cast<const float2 *>(gbp + cb + 8 + 2 * lc));
          }
          umma::tc_wait_ld();
          // y[i][j][u]: column cb + 8i + 2lc + u, row 32q + lr + 8j
          float y[2][4][2];
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const float2 sc = i ? a1 : a0, sh = i ? s1 : s0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                float acc = no_acc ? 0.f : __uint_as_float(v[j >> 1][4 * i + 2 * (j & 1) + u]);
                if (gbp) acc += u ? (i ? gb1.y : gb0.y) : (i ? gb1.x : gb0.x);
                if (xyz_term) {
                  const float4 w = s_wxyz[cb + 8 * i + 2 * lc + u];
                  acc = __fmaf_rn(w.x, rel[j].x, __fmaf_rn(w.y, rel[j].y, __fmaf_rn(w.z, rel[j].z, acc)));
                }
                float t = __fmaf_rn(u ? sc.y : sc.x, acc, u ? sh.y : sh.x);
                y[i][j][u] = t;
              }
          }
          if (!S.is_out) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const uint32_t sw = (uint32_t)(((cq * 2 + i) ^ lr) << 4);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                chain_store_act<MODE>(pan0 + pn * CH_PANEL + j * 1024 + sw, y[i][j][0], y[i][j][1], S.relu != 0);
            }
          } else if (gpool == 1) {
            // ---- dense output rows
            if (P.out_pool) {
              // max over the tile's rows per column (warp: 32 rows), merged across warps / tiles by the sign-aware atomic
              const int grp = l0 / P.pool_rows, ngrp = P.L / P.pool_rows;
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                float m0v = -INFINITY, m1v = -INFINITY;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  if (l0 + q * 32 + lr + 8 * j < P.L) {
                    float y0 = y[i][j][0], y1 = y[i][j][1];
                    if (S.relu) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); }
                    m0v = fmaxf(m0v, y0); m1v = fmaxf(m1v, y1);
                  }
                }
#pragma unroll
                for (int off = 4; off <= 16; off <<= 1) {
                  m0v = fmaxf(m0v, __shfl_xor_sync(0xffffffffu, m0v, off));
                  m1v = fmaxf(m1v, __shfl_xor_sync(0xffffffffu, m1v, off));
                }
                const int col = cb + 8 * i + 2 * lc;
                if (lr == 0 && m0v > -INFINITY) {
                  if (P.pool_pm) {
                    float *o = P.out_pool + ((size_t)b * ngrp + grp) * P.out_C + S.out_ch0 + col;
                    if (col < S.out_valid) atomic_max_float(o, m0v);
                    if (col + 1 < S.out_valid) atomic_max_float(o + 1, m1v);
                  } else {
                    float *o = P.out_pool + ((size_t)b * P.out_C + S.out_ch0 + col) * ngrp + grp;
                    if (col < S.out_valid) atomic_max_float(o, m0v);
                    if (col + 1 < S.out_valid) atomic_max_float(o + ngrp, m1v);
                  }
                }
              }
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int col = cb + 8 * i + 2 * lc;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int row = l0 + q * 32 + lr + 8 * j;
                if (row >= P.L) continue;
                float y0 = y[i][j][0], y1 = y[i][j][1];
                if (S.relu) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); }
                if (P.out_cm) {
                  float *o = P.out_cm + ((size_t)b * P.out_C + S.out_ch0 + col) * P.L + row;
                  if (col < S.out_valid) o[0] = y0;
                  if (col + 1 < S.out_valid) o[P.L] = y1;
                }
                if (P.out_pm) {
                  float *o = P.out_pm + ((size_t)b * P.L + row) * P.ldo + P.out_col0 + S.out_ch0 + col;
                  if (col < S.out_valid) o[0] = y0;
                  if (col + 1 < S.out_valid) o[1] = y1;
                }
              }
            }
          } else {
            // ---- ReLU (monotone: applied after the max), max over the rows of every group
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              // rows of this thread: 32q + lr + 8j.  pool=16: {j=0,1} and {j=2,3} are two groups; pool>=32: one.
              float g0[2], g1[2];
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                g0[u] = fmaxf(y[i][0][u], y[i][1][u]);
                g1[u] = fmaxf(y[i][2][u], y[i][3][u]);
                if (gpool >= 32) { g0[u] = fmaxf(g0[u], g1[u]); g1[u] = g0[u]; }
#pragma unroll
                for (int off = 4; off <= 16; off <<= 1) {
                  g0[u] = fmaxf(g0[u], __shfl_xor_sync(0xffffffffu, g0[u], off));
                  g1[u] = fmaxf(g1[u], __shfl_xor_sync(0xffffffffu, g1[u], off));
                }
                if (S.relu) { g0[u] = fmaxf(g0[u], 0.f); g1[u] = fmaxf(g1[u], 0.f); }
              }
              if (lr == 0) {
                const int col = cb + 8 * i + 2 * lc;
                const bool v0 = col < S.out_valid, v1 = col + 1 < S.out_valid;
                const int oc = S.out_ch0 + col;
                if (gpool == 16) {
                  const int ga = m0 + 2 * q, gb = ga + 1;
                  if (ga < Lo) {
                    if (P.out_cm) { float *o = P.out_cm + ((size_t)b * P.out_C + oc) * Lo + ga; if (v0) o[0] = g0[0]; if (v1) o[Lo] = g0[1]; }
                    if (P.out_pm) { float *o = P.out_pm + ((size_t)b * Lo + ga) * P.out_C + oc; if (v0) o[0] = g0[0]; if (v1) o[1] = g0[1]; }
                  }
                  if (gb < Lo) {
                    if (P.out_cm) { float *o = P.out_cm + ((size_t)b * P.out_C + oc) * Lo + gb; if (v0) o[0] = g1[0]; if (v1) o[Lo] = g1[1]; }
                    if (P.out_pm) { float *o = P.out_pm + ((size_t)b * Lo + gb) * P.out_C + oc; if (v0) o[0] = g1[0]; if (v1) o[1] = g1[1]; }
                  }
                } else if (gpool == 32) {
                  const int ga = m0 + q;
                  if (ga < Lo) {
                    if (P.out_cm) { float *o = P.out_cm + ((size_t)b * P.out_C + oc) * Lo + ga; if (v0) o[0] = g0[0]; if (v1) o[Lo] = g0[1]; }
                    if (P.out_pm) { float *o = P.out_pm + ((size_t)b * Lo + ga) * P.out_C + oc; if (v0) o[0] = g0[0]; if (v1) o[1] = g0[1]; }
                  }
                } else {
                  // pool 64 (two lane quarters per group) or 128 (four): outputs were zero-filled and the values are
                  // >= 0 after ReLU, so integer max == float max
                  const int ga = m0 + (q * 32) / gpool;
                  if (ga < Lo) {
                    if (P.out_cm) {
                      int *o = reinterpret_cast<int *>(P.out_cm + ((size_t)b * P.out_C + oc) * Lo + ga);
                      if (v0) atomicMax(o, __float_as_int(g0[0]));
                      if (v1) atomicMax(o + Lo, __float_as_int(g0[1]));
                    }
                    if (P.out_pm) {
                      int *o = reinterpret_cast<int *>(P.out_pm + ((size_t)b * Lo + ga) * P.out_C + oc);
                      if (v0) atomicMax(o, __float_as_int(g0[0]));
                      if (v1) atomicMax(o + 1, __float_as_int(g0[1]));
                    }
                  }
                }
              }
            }
          }
        }
        if (!S.is_out) {
          umma::fence_proxy_async_smem();
          umma::tc_fence_before();
          __syncwarp();
          if (lane == 0) umma::mbar_arrive(&bars->a_ready);
        } else {
          umma::tc_fence_before();
        }
      }
      // all epilogue warps are done with s_rel / the accumulators before the next tile's loader overwrites them
      asm volatile("bar.sync 1, %0;" ::"n"(32 * CH_EPI_WARPS) : "memory");
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    umma::tc_fence_after();
    umma::tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// wide_rows_kernel: ONE wide pointwise layer on row-major operands as a K-pipelined, role-split GEMM -- the shape the
// encoders of SkipPropagation need (262 144 rows, K up to 1536, 137 GFLOP per layer), where the chain kernel above (built
// for 0.5-4 GFLOP three-layer MLPs) loads, multiplies and stores a tile in sequence.
//   warp 0      weight producer: 4-slot ring of 32-KB bulk copies (two K panels of split weights in flight)
//   warp 1      MMA issuer: per 64-wide K panel 4 (fp16 / bf16) or 12 (x3) tcgen05.mma into one of TWO 256-column TMEM
//               accumulators (tile parity)
//   warps 2-9   loaders: fp32 rows -> (ReLU) -> 16-bit hi [+ lo] K-major swizzled A panel, 2-deep ring
//   warps 10-17 epilogue: TMEM -> scale * (acc + group bias) + shift (+ ReLU) -> row-major output and / or the sign-aware
//               max over row groups; it drains accumulator t while the loaders and the tensor pipe work on tile t + 1
// Weights / tables: the packed buffer of a single-layer rfd_mlp_chain_pack (same stage order and table offsets).
constexpr int WR_LOADERS = 16, WR_EPI = 8;
constexpr int WR_THREADS = 32 * (2 + WR_LOADERS + WR_EPI);  // 832
constexpr int WR_NSLOT = 4;
constexpr int WR_SM_A = 0;                               // [hi b0][hi b1][lo b0][lo b1], 16 KB each
constexpr int WR_SM_W = 4 * CH_PANEL;                    // 64 KB
constexpr int WR_SM_TAB = WR_SM_W + WR_NSLOT * CH_SLOT;  // scale[256] shift[256]
constexpr int WR_SM_BAR = WR_SM_TAB + 2 * 256 * 4;
constexpr int WR_SMEM_BYTES = WR_SM_BAR + 256 + 1024;
static_assert(WR_SMEM_BYTES <= 232448, "exceeds the 227 KB of dynamic shared memory per CTA");

struct WideBars {
  uint64_t w_full[WR_NSLOT], w_empty[WR_NSLOT];
  uint64_t a_full[2], a_empty[2];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

struct WideParams {
  const float *x;        // (R, ldi) row-major, operand = columns [0, K)
  int ldi, K, kp, R;
  const uint8_t *w;      // packed stages: per K panel hi [, lo] image of n x 128 B
  const float *scale, *shift;
  int n, n_valid, relu, relu_in;
  const float *gbias;    // (R / gbias_rows, n) or nullptr
  int gbias_rows;
  float *out;            // (R, ldo) or nullptr; columns [out_col0, out_col0 + n_valid)
  int ldo, out_col0;
  float *out_pool;       // (R / pool_rows, n_valid) or nullptr
  int pool_rows;
  int num_tiles;
};

template <int MODE>
__device__ __forceinline__ void wide_store_chunk(uint32_t dst, const float (&f)[8]) {
  uint32_t h[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    h[i] = MODE == CH_MODE_BF16 ? umma::pack_bf16x2(f[2 * i], f[2 * i + 1]) : umma::pack_f16x2(f[2 * i], f[2 * i + 1]);
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
  if (MODE == CH_MODE_F16X3) {
    uint32_t l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 hf = umma::unpack_f16x2(h[i]);
      l[i] = umma::pack_f16x2(f[2 * i] - hf.x, f[2 * i + 1] - hf.y);
    }
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 2 * CH_PANEL), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3])
                 : "memory");
  }
}

template <int MODE>
__global__ void __launch_bounds__(WR_THREADS, 1) wide_rows_kernel(const WideParams P) {
  constexpr bool X3 = MODE == CH_MODE_F16X3;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *s_a = smem + WR_SM_A, *s_w = smem + WR_SM_W;
  float *s_scale = reinterpret_cast<float *>(smem + WR_SM_TAB), *s_shift = s_scale + 256;
  WideBars *bars = reinterpret_cast<WideBars *>(smem + WR_SM_BAR);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int i = 0; i < WR_NSLOT; ++i) { umma::mbar_init(&bars->w_full[i], 1); umma::mbar_init(&bars->w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      umma::mbar_init(&bars->a_full[i], WR_LOADERS);
      umma::mbar_init(&bars->a_empty[i], 1);
      umma::mbar_init(&bars->acc_full[i], 1);
      umma::mbar_init(&bars->acc_empty[i], WR_EPI);
    }
    umma::fence_barrier_init();
  }
  if (warp == 1) umma::tmem_alloc(&bars->tmem_base, 512);
  for (int e = tid; e < 256; e += WR_THREADS) {
    s_scale[e] = e < P.n ? __ldg(P.scale + e) : 0.f;
    s_shift[e] = e < P.n ? __ldg(P.shift + e) : 0.f;
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const int tile_lo = (int)(((long long)P.num_tiles * blockIdx.x) / gridDim.x);
  const int tile_hi = (int)(((long long)P.num_tiles * (blockIdx.x + 1)) / gridDim.x);
  const uint32_t stage_bytes = (uint32_t)P.n * 128u;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t st = 0, ph = 0;
      for (int tile = tile_lo; tile < tile_hi; ++tile) {
        size_t off = 0;
        for (int i = 0; i < P.kp * (X3 ? 2 : 1); ++i) {
          umma::mbar_wait(&bars->w_empty[st], ph ^ 1u);
          umma::mbar_arrive_expect_tx(&bars->w_full[st], stage_bytes);
          umma::bulk_g2s(s_w + st * CH_SLOT, P.w + off, stage_bytes, &bars->w_full[st]);
          off += stage_bytes;
          if (++st == WR_NSLOT) { st = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t a_addr = umma::smem_u32(s_a), w_addr = umma::smem_u32(s_w);
      const uint32_t idesc = MODE == CH_MODE_BF16 ? umma::make_idesc_bf16_f32(CH_TILE_M, (uint32_t)P.n)
                                                  : umma::make_idesc_f16_f32(CH_TILE_M, (uint32_t)P.n);
      uint32_t st = 0, ph = 0, g = 0, t = 0;
      for (int tile = tile_lo; tile < tile_hi; ++tile, ++t) {
        const uint32_t acc = t & 1u;
        if (t >= 2) umma::mbar_wait(&bars->acc_empty[acc], ((t >> 1) - 1u) & 1u);  // the epilogue drained this accumulator
        umma::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256u;
        for (int kp = 0; kp < P.kp; ++kp, ++g) {
          const uint32_t b = g & 1u;
          umma::mbar_wait(&bars->a_full[b], (g >> 1) & 1u);
          umma::tc_fence_after();
          const uint32_t a_hi = a_addr + b * CH_PANEL, a_lo = a_hi + 2 * CH_PANEL;
          umma::mbar_wait(&bars->w_full[st], ph);
          umma::tc_fence_after();
          const uint32_t w_hi = w_addr + st * CH_SLOT;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma::mma_f16_ss(d_tmem, umma::make_desc_k_sw128(a_hi + k * 32), umma::make_desc_k_sw128(w_hi + k * 32), idesc,
                             (kp | k) ? 1u : 0u);
            if (X3)
              umma::mma_f16_ss(d_tmem, umma::make_desc_k_sw128(a_lo + k * 32), umma::make_desc_k_sw128(w_hi + k * 32), idesc, 1u);
          }
          umma::mma_commit(&bars->w_empty[st]);
          if (++st == WR_NSLOT) { st = 0; ph ^= 1u; }
          if (X3) {
            umma::mbar_wait(&bars->w_full[st], ph);
            umma::tc_fence_after();
            const uint32_t w_lo = w_addr + st * CH_SLOT;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma::mma_f16_ss(d_tmem, umma::make_desc_k_sw128(a_hi + k * 32), umma::make_desc_k_sw128(w_lo + k * 32), idesc, 1u);
            umma::mma_commit(&bars->w_empty[st]);
            if (++st == WR_NSLOT) { st = 0; ph ^= 1u; }
          }
          umma::mma_commit(&bars->a_empty[b]);
        }
        umma::mma_commit(&bars->acc_full[acc]);
      }
    }
  } else if (warp < 2 + WR_LOADERS) {
    // ---------------- loaders: lane = (row sub-index, 8-channel chunk); a warp instruction covers 4 rows x 64 channels
    const int wl = warp - 2, rsub = lane >> 3, ch = lane & 7;
    const uint32_t a_base = umma::smem_u32(s_a);
    constexpr int NIT = CH_TILE_M / (4 * WR_LOADERS);   // row groups per warp and panel (2)
    // software pipeline: the global loads of panel g + 1 are issued before panel g is converted and stored, so (with the
    // two-deep smem ring) up to three panels of a row tile are in flight per CTA -- the A operand comes from HBM at the
    // latency of a full memory round trip per panel otherwise
    float cur[NIT][8], nxt[NIT][8];
    auto fetch = [&](int tile, int kp, float (&f)[NIT][8]) {
      const long long l0 = (long long)tile * CH_TILE_M;
      const int c = kp * 64 + ch * 8;
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int row = 4 * (wl + WR_LOADERS * it) + rsub;
        const bool rv = l0 + row < P.R;
        const float *src = P.x + (size_t)(l0 + row) * P.ldi;
        if (rv && c + 8 <= P.K) {
          const float4 a = __ldg(reinterpret_cast<const float4 *>(src + c));
          const float4 d = __ldg(reinterpret_cast<const float4 *>(src + c + 4));
          f[it][0] = a.x; f[it][1] = a.y; f[it][2] = a.z; f[it][3] = a.w; f[it][4] = d.x; f[it][5] = d.y; f[it][6] = d.z; f[it][7] = d.w;
        } else {
#pragma unroll
          for (int u = 0; u < 8; ++u) f[it][u] = (rv && c + u < P.K) ? __ldg(src + c + u) : 0.f;
        }
      }
    };
    uint32_t g = 0;
    if (tile_lo < tile_hi) fetch(tile_lo, 0, cur);
    for (int tile = tile_lo; tile < tile_hi; ++tile) {
      for (int kp = 0; kp < P.kp; ++kp, ++g) {
        const uint32_t b = g & 1u;
        // next panel (of this tile, or the first of the next tile)
        const bool last_kp = kp + 1 == P.kp;
        const int ntile = last_kp ? tile + 1 : tile, nkp = last_kp ? 0 : kp + 1;
        if (ntile < tile_hi) fetch(ntile, nkp, nxt);
        if (g >= 2) umma::mbar_wait(&bars->a_empty[b], ((g >> 1) - 1u) & 1u);
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
          const int row = 4 * (wl + WR_LOADERS * it) + rsub;
          if (P.relu_in) {
#pragma unroll
            for (int u = 0; u < 8; ++u) cur[it][u] = fmaxf(cur[it][u], 0.f);
          }
          wide_store_chunk<MODE>(a_base + b * CH_PANEL + row * 128 + ((ch ^ (row & 7)) << 4), cur[it]);
        }
        umma::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(&bars->a_full[b]);
#pragma unroll
        for (int it = 0; it < NIT; ++it)
#pragma unroll
          for (int u = 0; u < 8; ++u) cur[it][u] = nxt[it][u];
      }
    }
  } else {
    // ---------------- epilogue warps (8): two per TMEM lane quarter, each owns two 16-column quarters of every panel
    const int we = warp - 2 - WR_LOADERS;      // 0..7
    const int q = warp & 3;                    // TMEM lane quarter = warp id % 4
    const int half = we >> 2;                  // which of the quarter's two warps
    const int lr = lane >> 2, lc = lane & 3;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int npan = (P.n + 63) >> 6;
    uint32_t t = 0;
    for (int tile = tile_lo; tile < tile_hi; ++tile, ++t) {
      const uint32_t acc = t & 1u;
      const long long l0 = (long long)tile * CH_TILE_M;
      umma::mbar_wait(&bars->acc_full[acc], (t >> 1) & 1u);
      umma::tc_fence_after();
      const uint32_t tmem_s = tmem_base + acc * 256u;
      const float *gbp = P.gbias ? P.gbias + (size_t)(l0 / P.gbias_rows) * P.n : nullptr;
#pragma unroll 1
      for (int pn = 0; pn < npan; ++pn) {
#pragma unroll 1
        for (int cqq = 0; cqq < 2; ++cqq) {
          const int cb = pn * 64 + (2 * half + cqq) * 16;
          if (cb >= P.n) continue;  // warp-uniform
          uint32_t v[2][8];
          umma::tmem_ld_16x256b_x2(tmem_s + lane_base + cb, v[0]);
          umma::tmem_ld_16x256b_x2(tmem_s + lane_base + (16u << 16) + cb, v[1]);
          float2 gb[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
          if (gbp) {
            gb[0] = __ldg(reinterpret_cast<const float2 *>(gbp + cb + 2 * lc));
            gb[1] = __ldg(reinterpret_cast<const float2 *>(gbp + cb + 8 + 2 * lc));
          }
          umma::tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int col = cb + 8 * i + 2 * lc;
            const float sc0 = s_scale[col], sc1 = s_scale[col + 1], sh0 = s_shift[col], sh1 = s_shift[col + 1];
            float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const long long row = l0 + q * 32 + lr + 8 * j;
              float y0 = __fmaf_rn(sc0, __uint_as_float(v[j >> 1][4 * i + 2 * (j & 1)]) + gb[i].x, sh0);
              float y1 = __fmaf_rn(sc1, __uint_as_float(v[j >> 1][4 * i + 2 * (j & 1) + 1]) + gb[i].y, sh1);
              if (P.relu) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); }
              if (row < P.R) {
                m0 = fmaxf(m0, y0); m1 = fmaxf(m1, y1);
                if (P.out) {
                  float *o = P.out + (size_t)row * P.ldo + P.out_col0 + col;
                  if (col < P.n_valid) o[0] = y0;
                  if (col + 1 < P.n_valid) o[1] = y1;
                }
              }
            }
            if (P.out_pool) {
#pragma unroll
              for (int off = 4; off <= 16; off <<= 1) {
                m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, off));
                m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, off));
              }
              if (lr == 0 && m0 > -INFINITY) {
                float *o = P.out_pool + (size_t)(l0 / P.pool_rows) * P.n_valid + col;
                if (col < P.n_valid) atomic_max_float(o, m0);
                if (col + 1 < P.n_valid) atomic_max_float(o + 1, m1);
              }
            }
          }
        }
      }
      umma::tc_fence_before();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(&bars->acc_empty[acc]);
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    umma::tc_fence_after();
    umma::tmem_dealloc(tmem_base, 512);
  }
}

// pack one weight stage set: W (n_valid rows, ld columns) f32 row-major, columns [col0, col0 + K) are the K operand;
// rows [row0, row0 + n_valid) of W map to stage rows 0..n_valid-1, rows up to n_pad are zero.
// -> for every k panel: hi image [n_pad rows][64 k] K-major 128B-swizzled (then the lo image in x3 mode)
__global__ void chain_pack_kernel(const float *__restrict__ W, int ld, int col0, int K, int row0, int n_valid, int n_pad,
                                  int kpn, int mode, uint8_t *__restrict__ dst) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;  // one thread per 16-byte chunk
  const int total = kpn * n_pad * 8;
  if (e >= total) return;
  const int kp = e / (n_pad * 8), rem = e % (n_pad * 8);
  const int n = rem / 8, cin = rem % 8;
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k0 = kp * 64 + cin * 8 + 2 * i;
    const float a = (n < n_valid && k0 < K) ? __ldg(W + (size_t)(row0 + n) * ld + col0 + k0) : 0.f;
    const float b = (n < n_valid && k0 + 1 < K) ? __ldg(W + (size_t)(row0 + n) * ld + col0 + k0 + 1) : 0.f;
    if (mode == CH_MODE_BF16) {
      hi[i] = umma::pack_bf16x2(a, b);
      lo[i] = 0;
    } else {
      hi[i] = umma::pack_f16x2(a, b);
      const float2 hf = umma::unpack_f16x2(hi[i]);
      lo[i] = umma::pack_f16x2(a - hf.x, b - hf.y);
    }
  }
  const size_t stage = (size_t)n_pad * 128;
  const size_t per_kp = stage * (mode == CH_MODE_F16X3 ? 2 : 1);
  uint8_t *d = dst + (size_t)kp * per_kp + n * 128 + ((cin ^ (n & 7)) << 4);
  *reinterpret_cast<uint4 *>(d) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  if (mode == CH_MODE_F16X3) *reinterpret_cast<uint4 *>(d + stage) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

__global__ void chain_pack_xyz_kernel(const float *__restrict__ W, int ld, int n_valid, int n_pad, float4 *__restrict__ dst) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_pad) return;
  dst[n] = n < n_valid ? make_float4(__ldg(W + (size_t)n * ld), __ldg(W + (size_t)n * ld + 1), __ldg(W + (size_t)n * ld + 2), 0.f)
                       : make_float4(0.f, 0.f, 0.f, 0.f);
}

// ---- host-side plan shared by packed_bytes / pack / run
struct ChainPlan {
  int nsteps = 0;
  ChainStep st[CH_MAX_STEPS];
  int k_valid[CH_MAX_STEPS];   // valid K of each step
  int layer[CH_MAX_STEPS];     // source layer of each step
  int row0[CH_MAX_STEPS];      // first weight row of the step inside its layer
  size_t w_bytes = 0;          // packed weight stages
  int tab_floats = 0;
  int out_C = 0;
  bool ok = false;
};

static ChainPlan chain_plan(int mode, int K0, int xyz, int C1, int C2, int C3) {
  ChainPlan p;
  if (mode < CH_MODE_BF16 || mode > CH_MODE_F16X3 || K0 < 0 || (xyz != 0 && xyz != 3) || C1 < 1) return p;
  if (K0 == 0 && !xyz) return p;
  const int widths[3] = {C1, C2, C3};
  const int nl = C3 > 0 ? 3 : (C2 > 0 ? 2 : 1);
  if (nl >= 2 && C2 < 1) return p;
  int k = K0;
  int tab = 0;
  for (int l = 0; l < nl; ++l) {
    const bool last = l == nl - 1;
    const int n = widths[l];
    if (!last) {
      if (n > 256) return p;
      ChainStep &s = p.st[p.nsteps];
      s = ChainStep{(k + 63) / 64, (n + 63) & ~63, 1, 0, 0, 0, n, tab};
      p.k_valid[p.nsteps] = k; p.layer[p.nsteps] = l; p.row0[p.nsteps] = 0;
      tab += s.n;
      ++p.nsteps;
      k = n;
    } else {
      if (n > 512) return p;
      for (int c0 = 0; c0 < n; c0 += 256) {
        if (p.nsteps >= CH_MAX_STEPS) return p;
        const int nv = n - c0 < 256 ? n - c0 : 256;
        // a second column block re-reads the RESIDENT A panels: its K must fit in them (layer 0 with K0 > 256 streams
        // through the panels in rounds and is gone by then)
        if (c0 > 0 && (k + 63) / 64 > CH_APAN) return p;
        ChainStep &s = p.st[p.nsteps];
        s = ChainStep{(k + 63) / 64, (nv + 15) & ~15, 1, 1, c0 > 0 ? 1 : 0, c0, nv, tab};
        p.k_valid[p.nsteps] = k; p.layer[p.nsteps] = l; p.row0[p.nsteps] = c0;
        tab += (s.n + 63) & ~63;  // the epilogue reads the tables in 16-column quarters of 64-column panels
        ++p.nsteps;
      }
      p.out_C = n;
    }
  }
  if (p.st[0].kp == 0 && !xyz) return p;
  if (tab > CH_TAB) return p;
  p.tab_floats = tab;
  for (int s = 0; s < p.nsteps; ++s) p.w_bytes += (size_t)p.st[s].kp * p.st[s].n * 128 * (mode == CH_MODE_F16X3 ? 2 : 1);
  p.ok = true;
  return p;
}

}  // namespace rfd

using namespace rfd;

// packed buffer layout: [weight stages (w_bytes, 256-aligned)][w_xyz: 256 x float4][scale: CH_TAB f32][shift: CH_TAB f32]
static size_t chain_off_xyz(const ChainPlan &p) { return (p.w_bytes + 255) & ~(size_t)255; }
static size_t chain_off_scale(const ChainPlan &p) { return chain_off_xyz(p) + 256 * 16; }
static size_t chain_off_shift(const ChainPlan &p) { return chain_off_scale(p) + CH_TAB * 4; }

extern "C" size_t rfd_mlp_chain_packed_bytes(int mode, int K0, int xyz, int C1, int C2, int C3) {
  const ChainPlan p = chain_plan(mode, K0, xyz, C1, C2, C3);
  return p.ok ? chain_off_shift(p) + CH_TAB * 4 : 0;
}

extern "C" int rfd_mlp_chain_pack(int mode, int K0, int xyz, const float *W1, const float *scale1, const float *shift1,
                                  int C1, const float *W2, const float *scale2, const float *shift2, int C2,
                                  const float *W3, const float *scale3, const float *shift3, int C3, int relu_last,
                                  void *packed, void *stream) {
  const ChainPlan p = chain_plan(mode, K0, xyz, C1, C2, C3);
  if (!p.ok) return RFD_ERR_UNSUPPORTED_SIZE;
  if (!W1 || !scale1 || !shift1 || !packed || (C2 > 0 && (!W2 || !scale2 || !shift2)) || (C3 > 0 && (!W3 || !scale3 || !shift3)))
    return RFD_ERR_INVALID_ARGUMENT;
  cudaStream_t st = as_stream(stream);
  uint8_t *base = reinterpret_cast<uint8_t *>(packed);
  const float *Ws[3] = {W1, W2, W3}, *Sc[3] = {scale1, scale2, scale3}, *Sh[3] = {shift1, shift2, shift3};
  const int widths[3] = {C1, C2, C3};
  RFD_CHECK_CUDA(cudaMemsetAsync(base + chain_off_xyz(p), 0, 256 * 16 + 2 * CH_TAB * 4, st), "mlp_chain_pack memset");
  uint8_t *dst = base;
  for (int s = 0; s < p.nsteps; ++s) {
    const int l = p.layer[s];
    const int ld = l == 0 ? xyz + K0 : widths[l - 1];
    const int col0 = l == 0 ? xyz : 0;
    const int kpn = p.st[s].kp;
    const int n_valid = p.st[s].is_out ? p.st[s].out_valid : widths[l];
    if (kpn > 0) {
      const int total = kpn * p.st[s].n * 8;
      chain_pack_kernel<<<h_ceil_div(total, 256), 256, 0, st>>>(Ws[l], ld, col0, p.k_valid[s], p.row0[s], n_valid, p.st[s].n,
                                                               kpn, mode, dst);
      RFD_CHECK_LAUNCH("chain_pack_kernel");
    }
    dst += (size_t)kpn * p.st[s].n * 128 * (mode == CH_MODE_F16X3 ? 2 : 1);
    RFD_CHECK_CUDA(cudaMemcpyAsync(base + chain_off_scale(p) + (size_t)p.st[s].tab_off * 4, Sc[l] + p.row0[s],
                                   (size_t)n_valid * 4, cudaMemcpyDeviceToDevice, st), "mlp_chain_pack scale");
    RFD_CHECK_CUDA(cudaMemcpyAsync(base + chain_off_shift(p) + (size_t)p.st[s].tab_off * 4, Sh[l] + p.row0[s],
                                   (size_t)n_valid * 4, cudaMemcpyDeviceToDevice, st), "mlp_chain_pack shift");
  }
  if (xyz) {
    chain_pack_xyz_kernel<<<1, 256, 0, st>>>(W1, xyz + K0, C1, p.st[0].n, reinterpret_cast<float4 *>(base + chain_off_xyz(p)));
    RFD_CHECK_LAUNCH("chain_pack_xyz_kernel");
  }
  (void)relu_last;
  return RFD_OK;
}

static int chain_launch(int mode, ChainParams &P, const ChainPlan &p, const void *packed, int relu_last, int B, int L,
                        int pool, float *out_cm, float *out_pm, void *stream) {
  const uint8_t *base = reinterpret_cast<const uint8_t *>(packed);
  P.w = base;
  P.w_xyz = reinterpret_cast<const float *>(base + chain_off_xyz(p));
  P.scale = reinterpret_cast<const float *>(base + chain_off_scale(p));
  P.shift = reinterpret_cast<const float *>(base + chain_off_shift(p));
  P.tab_floats = p.tab_floats;
  P.out_cm = out_cm; P.out_pm = out_pm; P.out_C = p.out_C;
  if (P.ldo == 0) P.ldo = p.out_C;
  P.B = B; P.L = L; P.pool = pool;
  P.nsteps = p.nsteps;
  for (int s = 0; s < p.nsteps; ++s) {
    P.st[s] = p.st[s];
    if (p.st[s].is_out) P.st[s].relu = relu_last;
  }
  P.tiles_per_scene = (L + CH_TILE_M - 1) / CH_TILE_M;
  const long long nt = (long long)P.tiles_per_scene * B;
  if (nt > 0x7fffffffLL) return RFD_ERR_UNSUPPORTED_SIZE;
  P.num_tiles = (int)nt;
  cudaStream_t st = as_stream(stream);
  if (pool > 32) {
    if (!relu_last) return RFD_ERR_UNSUPPORTED_SIZE;  // the atomicMax merge needs non-negative values
    const size_t n = sizeof(float) * (size_t)B * p.out_C * (L / pool);
    if (out_cm) RFD_CHECK_CUDA(cudaMemsetAsync(out_cm, 0, n, st), "mlp_chain memset");
    if (out_pm) RFD_CHECK_CUDA(cudaMemsetAsync(out_pm, 0, n, st), "mlp_chain memset");
  }
  int dev = 0, sms = 148;
  RFD_CHECK_CUDA(cudaGetDevice(&dev), "mlp_chain getdevice");
  RFD_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "mlp_chain sms");
  const int grid = (int)(nt < sms ? nt : sms);
#define RFD_CHAIN_LAUNCH(MODE)                                                                                      \
  do {                                                                                                              \
    RFD_CHECK_CUDA(cudaFuncSetAttribute(mlp_chain_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                        CH_SMEM_BYTES), "mlp_chain attr");                                          \
    mlp_chain_tc_kernel<MODE><<<grid, CH_THREADS, CH_SMEM_BYTES, st>>>(P);                                          \
  } while (0)
  if (mode == CH_MODE_BF16) RFD_CHAIN_LAUNCH(CH_MODE_BF16);
  else if (mode == CH_MODE_F16) RFD_CHAIN_LAUNCH(CH_MODE_F16);
  else RFD_CHAIN_LAUNCH(CH_MODE_F16X3);
#undef RFD_CHAIN_LAUNCH
  RFD_CHECK_LAUNCH("mlp_chain_tc_kernel");
  return RFD_OK;
}

extern "C" int rfd_mlp_chain_ex(int mode, const float *x, int B, int K0, int L, const void *packed, int C1, int C2, int C3,
                                int relu_last, int pool, float *out_cm, float *out_pm, int relu_in, const float *gbias,
                                int gbias_rows, float *out_pool, int pool_rows, void *stream) {
  if (B < 0 || L < 0 || K0 < 1 || pool < 1) return RFD_ERR_INVALID_ARGUMENT;
  if (B == 0 || L == 0) return RFD_OK;
  if (!x || !packed || (!out_cm && !out_pm && !out_pool)) return RFD_ERR_INVALID_ARGUMENT;
  if (gbias && (gbias_rows < CH_TILE_M || gbias_rows % CH_TILE_M || L % gbias_rows)) return RFD_ERR_UNSUPPORTED_SIZE;
  if (out_pool && (pool != 1 || pool_rows < CH_TILE_M || pool_rows % CH_TILE_M || L % pool_rows)) return RFD_ERR_UNSUPPORTED_SIZE;
  // pool: 1 (none), 16 / 32 (inside a warp's rows), or any power of two >= 64 (merged across warps / tiles with atomicMax)
  if (!(pool == 1 || pool == 16 || pool == 32 || (pool >= 64 && (pool & (pool - 1)) == 0)) || L % pool) return RFD_ERR_UNSUPPORTED_SIZE;
  const ChainPlan p = chain_plan(mode, K0, 0, C1, C2, C3);
  if (!p.ok) return RFD_ERR_UNSUPPORTED_SIZE;
  ChainParams P = {};
  P.x = x; P.K0 = K0; P.M = L / pool; P.S = pool;
  P.relu_in = relu_in; P.gbias = gbias; P.gbias_rows = gbias_rows; P.out_pool = out_pool; P.pool_rows = pool_rows;
  return chain_launch(mode, P, p, packed, relu_last, B, L, pool, out_cm, out_pm, stream);
}

extern "C" int rfd_mlp_chain_rows(int mode, const float *x_pm, int ldi, int B, int K0, int L, const void *packed, int C1, int C2,
                                  int C3, int relu_last, float *out_pm, int ldo, int out_col0, int relu_in,
                                  const float *gbias, int gbias_rows, float *out_pool, int pool_rows, void *stream) {
  if (B < 0 || L < 0 || K0 < 1 || ldi < K0 || ldo < 0 || out_col0 < 0) return RFD_ERR_INVALID_ARGUMENT;
  if (B == 0 || L == 0) return RFD_OK;
  if (!x_pm || !packed || (!out_pm && !out_pool)) return RFD_ERR_INVALID_ARGUMENT;
  if ((ldi & 3) || (reinterpret_cast<uintptr_t>(x_pm) & 15)) return RFD_ERR_INVALID_ARGUMENT;  // 16-byte row loads
  if (gbias && (gbias_rows < CH_TILE_M || gbias_rows % CH_TILE_M || L % gbias_rows)) return RFD_ERR_UNSUPPORTED_SIZE;
  if (out_pool && (pool_rows < CH_TILE_M || pool_rows % CH_TILE_M || L % pool_rows)) return RFD_ERR_UNSUPPORTED_SIZE;
  const ChainPlan p = chain_plan(mode, K0, 0, C1, C2, C3);
  if (!p.ok) return RFD_ERR_UNSUPPORTED_SIZE;
  if (out_pm && ldo < out_col0 + p.out_C) return RFD_ERR_INVALID_ARGUMENT;
  if (p.nsteps == 1) {
    // one layer (<= 256 outputs): the K-pipelined, role-split kernel; (B, L) rows are one flat list of B*L rows
    const long long R = (long long)B * L, nt = (R + CH_TILE_M - 1) / CH_TILE_M;
    if (nt > 0x7fffffffLL) return RFD_ERR_UNSUPPORTED_SIZE;
    const uint8_t *base = reinterpret_cast<const uint8_t *>(packed);
    WideParams W = {};
    W.x = x_pm; W.ldi = ldi; W.K = K0; W.kp = p.st[0].kp; W.R = (int)R;
    W.w = base;
    W.scale = reinterpret_cast<const float *>(base + chain_off_scale(p));
    W.shift = reinterpret_cast<const float *>(base + chain_off_shift(p));
    W.n = p.st[0].n; W.n_valid = p.st[0].out_valid; W.relu = relu_last; W.relu_in = relu_in;
    W.gbias = gbias; W.gbias_rows = gbias_rows;
    W.out = out_pm; W.ldo = ldo; W.out_col0 = out_col0;
    W.out_pool = out_pool; W.pool_rows = pool_rows;
    W.num_tiles = (int)nt;
    int dev = 0, sms = 148;
    RFD_CHECK_CUDA(cudaGetDevice(&dev), "wide_rows getdevice");
    RFD_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "wide_rows sms");
    const int grid = (int)(nt < sms ? nt : sms);
    cudaStream_t st = as_stream(stream);
#define RFD_WIDE_LAUNCH(MODE)                                                                                       \
  do {                                                                                                              \
    RFD_CHECK_CUDA(cudaFuncSetAttribute(wide_rows_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                        WR_SMEM_BYTES), "wide_rows attr");                                          \
    wide_rows_kernel<MODE><<<grid, WR_THREADS, WR_SMEM_BYTES, st>>>(W);                                             \
  } while (0)
    if (mode == CH_MODE_BF16) RFD_WIDE_LAUNCH(CH_MODE_BF16);
    else if (mode == CH_MODE_F16) RFD_WIDE_LAUNCH(CH_MODE_F16);
    else RFD_WIDE_LAUNCH(CH_MODE_F16X3);
#undef RFD_WIDE_LAUNCH
    RFD_CHECK_LAUNCH("wide_rows_kernel");
    return RFD_OK;
  }
  ChainParams P = {};
  P.x_pm = x_pm; P.ldi = ldi; P.K0 = K0; P.M = L; P.S = 1;
  P.ldo = out_pm ? ldo : p.out_C; P.out_col0 = out_col0; P.pool_pm = 1;
  P.relu_in = relu_in; P.gbias = gbias; P.gbias_rows = gbias_rows; P.out_pool = out_pool; P.pool_rows = pool_rows;
  return chain_launch(mode, P, p, packed, relu_last, B, L, 1, nullptr, out_pm, stream);
}

extern "C" int rfd_mlp_chain(int mode, const float *x, int B, int K0, int L, const void *packed, int C1, int C2, int C3,
                             int relu_last, int pool, float *out_cm, float *out_pm, void *stream) {
  if (!out_cm && !out_pm) return (B == 0 || L == 0) ? RFD_OK : RFD_ERR_INVALID_ARGUMENT;
  return rfd_mlp_chain_ex(mode, x, B, K0, L, packed, C1, C2, C3, relu_last, pool, out_cm, out_pm, 0, nullptr, 0, nullptr, 0,
                          stream);
}

extern "C" int rfd_sa_mlp_chain(int mode, const float *xyz, const float *new_xyz, const float *feat_pm, const int *idx, int B,
                                int N, int M, int S, int C, float radius, int normalize_xyz, const void *packed, int C1,
                                int C2, int C3, float *out_cm, float *out_pm, void *stream) {
  if (B < 0 || M < 0 || S < 1 || N < 1 || C < 0) return RFD_ERR_INVALID_ARGUMENT;
  if (B == 0 || M == 0) return RFD_OK;
  if (!xyz || !new_xyz || !idx || (C > 0 && !feat_pm) || !packed || (!out_cm && !out_pm)) return RFD_ERR_INVALID_ARGUMENT;
  if (!(S == 16 || S == 32 || S == 64 || S == 128)) return RFD_ERR_UNSUPPORTED_SIZE;
  const ChainPlan p = chain_plan(mode, C, 3, C1, C2, C3);
  if (!p.ok) return RFD_ERR_UNSUPPORTED_SIZE;
  ChainParams P = {};
  P.idx = idx; P.xyz = xyz; P.new_xyz = new_xyz; P.feat_pm = feat_pm; P.N = N;
  P.inv_r = normalize_xyz ? 1.0f / radius : 1.0f;
  P.normalize = normalize_xyz ? 1 : 0;
  P.K0 = C; P.has_xyz = 1; P.M = M; P.S = S;
  return chain_launch(mode, P, p, packed, 1, B, M * S, S, out_cm, out_pm, stream);
}
