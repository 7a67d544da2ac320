// onet_decoder.cu -- the ONet occupancy decoder (DecoderCBatchNorm, eval mode) for sm_100a.
//
// Reference: /root/reference/models/iscnet/modules/occ_decoder.py:72-122 (DecoderCBatchNorm.forward :110-122),
// layers.py:51-107 (CResnetBlockConv1d.forward :98-107), layers.py:193-242 (CBatchNorm1d.forward :226-242);
// driver generator.py:123-143 (eval_points, ONE object of 32768 points per call, 11 conv launches + 22 CBN
// convs + elementwise kernels per object, then a .cpu() sync).
//
//   net = fc_p(p) + fc_z(z)                                   (3 -> 256)
//   5 x { net = net + fc_1(relu(cbn_1(fc_0(relu(cbn_0(net, c))), c))) }   (256 -> 256 -> 256)
//   out = fc_out(relu(cbn(net, c)))                            (256 -> 1)
//   cbn_l(x, c)[ch] = gamma_l(c)[ch] * (x[ch] - mean_l[ch]) / sqrt(var_l[ch] + eps) + beta_l(c)[ch]   (eval mode)
//
// Here the whole decoder for a tile of 128 query points is ONE pass of a persistent, warp-specialised kernel:
//   * the fp32 residual stream x (128 x 256) and the hidden pre-activation net (128 x 256) live in TENSOR MEMORY
//     (2 x 256 columns = the whole 512-column TMEM of the SM); fc_1 accumulates straight onto x, so the residual
//     add costs nothing and never leaves fp32;
//   * the ten 256x256 GEMMs run on tcgen05.mma (M=128, N=256, K=16, bf16 x bf16 -> f32), issued by one thread;
//   * conv biases are never added in the kernel: they are folded, together with the conditional-BN statistics,
//     into one per-object/per-layer/per-channel affine (a, c) so that relu(cbn_l(x_true)) = relu(a*x_acc + c)
//     (rfd_onet_cbn_tables).  The epilogue warps read the accumulator with tcgen05.ld, apply that affine + ReLU,
//     round to bf16 and write the next layer's A operand into shared memory in the UMMA K-major 128B-swizzled
//     layout, 64-channel panel by panel; the MMA of the next layer starts on a panel as soon as it is complete;
//   * weights (10 x 256 x 256 bf16, pre-swizzled by rfd_onet_pack_weights) stream from L2 through a 4-stage ring
//     of 32-KB bulk copies (cp.async.bulk, mbarrier complete_tx) issued by a dedicated producer warp; each CTA owns a
//     contiguous chunk of tiles, so the per-object affine table is reloaded only when the object changes;
//   * tried, correct, but no faster (profiles/r1_decoder_history.md): TS-mode fc_1 -- the epilogue writes
//     relu(cbn_1(net)) as packed bf16 back INTO tensor memory over net's own columns (tcgen05.st.16x128b) and the MMA
//     reads A from TMEM, removing half of the A-panel shared-memory stores, the proxy fences and the A-operand
//     shared-memory reads.  8.96 ms vs 8.67 ms: the layer time is a latency chain, not shared-memory bandwidth
//     (the primitives and their self-test stay: umma::mma_bf16_ts, rfd_umma_selftest_ts);
//   * tried and rejected (profiles/r1_decoder_history.md): issuing each layer as two N=128 halves so that the
//     epilogue of the first half hides behind the MMAs of the second -- the A panels would have to be double
//     buffered (WAR hazard on the in-place activation panels) and the N=128 MMAs re-read A twice, which made the
//     kernel shared-memory bound (13.7 ms vs 8.7 ms for 256 objects);
//   * fc_p (K = 3) and fc_out (N = 1) are evaluated in fp32 on the CUDA cores inside the same kernel.
#include "common.cuh"
#include "umma.cuh"
#include <atomic>
#include <mutex>
#include <cstdlib>

namespace rfd {

constexpr int DEC_H = 256;
constexpr int DEC_LAYERS = 10;
constexpr int DEC_CBN = 11;
constexpr int DEC_TILE_M = 128;
constexpr int DEC_KP = 4;
constexpr int DEC_PANEL_A = DEC_TILE_M * 128;   // 16384 B : 128 rows x 64 16-bit elements
constexpr int DEC_STAGE_B = DEC_H * 128;        // 32768 B : 256 rows x 64 16-bit elements
// operand modes (rfdnet_b200.h RFD_ONET_MODE_*)
constexpr int MODE_BF16 = 1;    // bf16 x bf16, one MMA per K step
constexpr int MODE_F16 = 2;     // fp16 x fp16 (11 significant bits instead of 8), one MMA per K step
constexpr int MODE_F16X3 = 3;   // split fp16: a_hi.w_hi + a_lo.w_hi + a_hi.w_lo  (~22 significant bits)
// per-object record in global memory (rfd_onet_cbn_tables):
//   [0, 5632)      plain   [11 layers][a: 256][c: 256]                      (fp32 path)
//   [5632, 5888)   x_bias  [256]
//   [5888, 11520)  paired  [11 layers][128 column pairs]{a0, a1, c0, c1}    (tensor-core path)
// the tensor-core kernel stages [5632, 11520) = DEC_AFF_FLOATS floats per object.
constexpr int DEC_PLAIN_FLOATS = DEC_CBN * 2 * DEC_H;         // 5632
constexpr int DEC_AFF_FLOATS = DEC_H + DEC_CBN * 2 * DEC_H;   // 5888 floats staged in shared memory
constexpr int DEC_AFF_BYTES = DEC_AFF_FLOATS * 4;             // 23552
constexpr int DEC_REC_FLOATS = DEC_PLAIN_FLOATS + DEC_AFF_FLOATS;  // 11520
constexpr int DEC_EPI_WARPS = 16;               // 4 TMEM lane quarters x 4 column quarters of every 64-column panel
constexpr int DEC_THREADS = 64 + 32 * DEC_EPI_WARPS;  // warp 0 producer, warp 1 MMA, warps 2.. epilogue
constexpr int DEC_CW = 16;                      // columns of a panel owned by one epilogue warp
constexpr int DEC_MAX_STAGES = 4;

// shared memory map (offsets from a 1024-B aligned base).  Activation panels + weight ring always fill 192 KB:
//   one-MMA modes : A 64 KB                | 4-stage weight ring 128 KB
//   MODE_F16X3    : A_hi 64 KB, A_lo 64 KB | 2-stage weight ring  64 KB
constexpr int SM_AH = 0;
constexpr int SM_AL = DEC_KP * DEC_PANEL_A;                     // MODE_F16X3 only
constexpr int SM_AFF = 6 * DEC_STAGE_B;                         // 196608
constexpr int SM_WP = SM_AFF + DEC_AFF_BYTES;                   // single affine buffer (reloaded on object change)
constexpr int SM_WOUT = SM_WP + 3 * DEC_H * 4;
constexpr int SM_OUT = SM_WOUT + DEC_H * 4;
constexpr int SM_BAR = SM_OUT + 4 * DEC_TILE_M * 4;
constexpr int SM_TOTAL = SM_BAR + 256;
constexpr int DEC_SMEM_BYTES = SM_TOTAL + 1024;                 // + alignment slack
static_assert(DEC_SMEM_BYTES <= 232448, "exceeds the 227 KB of dynamic shared memory per CTA");

struct DecBars {
  uint64_t w_full[DEC_MAX_STAGES];
  uint64_t w_empty[DEC_MAX_STAGES];
  uint64_t acc_ready;
  uint32_t tmem_base;
};

// epilogue value pair -> operand panel(s): relu, round, store.  MODE_F16X3 also stores the rounding residual
// (exact in fp32: lo = relu(v) - float(hi)) into the lo panel 64 KB above.
template <int MODE>
__device__ __forceinline__ void store_act_pair(uint32_t addr, float v0, float v1) {
  if (MODE == MODE_BF16) {
    umma::sts_u32(addr, umma::pack_relu_bf16x2(v0, v1));
  } else {
    const uint32_t h = umma::pack_relu_f16x2(v0, v1);
    umma::sts_u32(addr, h);
    if (MODE == MODE_F16X3) {
      const float2 hf = umma::unpack_f16x2(h);
      umma::sts_u32(addr + SM_AL, umma::pack_f16x2(fmaxf(v0, 0.f) - hf.x, fmaxf(v1, 0.f) - hf.y));
    }
  }
}

// TRACE: CTA 0 records clock64() at the hand-off points of its first two tiles (diagnostics, rfd_onet_decode_traced):
//   trace[tile][layer][0..3] MMA thread: operands of panel kp ready ; [4..7] MMAs of panel kp issued + committed
//   trace[tile][layer][8]    epilogue warp 2: acc_ready observed     ; [9..12] panel kp published (named barrier arrive)
//
// CL (cluster size 1 or 2): with CL = 2 the two CTAs of a cluster stream the SAME weight stages; each CTA fetches one
// 16-KB half of every stage from L2 and MULTICASTS it into both CTAs' rings (cp.async.bulk ... .multicast::cluster), so
// every weight byte crosses the L2 -> SM fabric once per pair instead of once per CTA.  A ring slot is refilled only
// when both CTAs' MMAs have retired it (tcgen05.commit ... .multicast::cluster arrives on both CTAs' w_empty barrier,
// which therefore counts CL arrivals).  Everything else (TMEM, activation panels, hand-offs) stays CTA-local.
template <int MODE, int CL, bool TRACE>
__global__ void __launch_bounds__(DEC_THREADS, 1)
onet_decode_kernel(const float *__restrict__ p, long long p_stride, int T, const float *__restrict__ fc_p_w,
                   const uint8_t *__restrict__ packed, const float *__restrict__ aff_all,
                   const float *__restrict__ fc_out_w, float fc_out_b, float *__restrict__ logits, int num_tiles,
                   int tiles_per_obj, unsigned long long *__restrict__ trace) {
  constexpr int NSTAGE = MODE == MODE_F16X3 ? 2 : 4;
  constexpr int SPK = MODE == MODE_F16X3 ? 2 : 1;  // weight stages per (layer, k-panel): hi [, lo]
  constexpr int SM_W = MODE == MODE_F16X3 ? 2 * DEC_KP * DEC_PANEL_A : DEC_KP * DEC_PANEL_A;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *s_ah = smem + SM_AH;
  uint8_t *s_w = smem + SM_W;
  float *s_aff = reinterpret_cast<float *>(smem + SM_AFF);
  float *s_wp = reinterpret_cast<float *>(smem + SM_WP);      // [3][256]
  float *s_wout = reinterpret_cast<float *>(smem + SM_WOUT);  // [256]
  float *s_out = reinterpret_cast<float *>(smem + SM_OUT);    // [4][128]
  DecBars *bars = reinterpret_cast<DecBars *>(smem + SM_BAR);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t crank = CL > 1 ? umma::cluster_ctarank() : 0u;

  if (tid == 0) {
    for (int i = 0; i < NSTAGE; ++i) { umma::mbar_init(&bars->w_full[i], 1); umma::mbar_init(&bars->w_empty[i], CL); }
    umma::mbar_init(&bars->acc_ready, 1);
    umma::fence_barrier_init();
  }
  if (warp == 1) umma::tmem_alloc(&bars->tmem_base, 512);
  for (int e = tid; e < 3 * DEC_H; e += DEC_THREADS) {
    const int k = e / DEC_H, o = e % DEC_H;
    s_wp[e] = __ldg(fc_p_w + o * 3 + k);  // torch layout (256,3) -> [3][256]
  }
  for (int e = tid; e < DEC_H; e += DEC_THREADS) s_wout[e] = __ldg(fc_out_w + e);
  umma::tc_fence_before();
  __syncthreads();
  if (CL > 1) umma::cluster_sync();  // the peer's barriers exist before anything is multicast into this CTA
  umma::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const uint32_t tmem_x = tmem_base, tmem_n = tmem_base + DEC_H;
  // contiguous chunk of tiles per cluster, dealt round-robin to its CTAs: consecutive tiles of a CTA almost always belong
  // to the same object, so the per-object affine table is (re)loaded only on an object change (2-3 times per launch).
  // Every CTA of a cluster runs the same number of weight passes (n_slots); a CTA without a tile in the last slot
  // only drains the ring (phantom pass).
  const int ncl = gridDim.x / CL, cid = blockIdx.x / CL;
  const int c_lo = (int)(((long long)num_tiles * cid) / ncl);
  const int c_hi = (int)(((long long)num_tiles * (cid + 1)) / ncl);
  const int n_slots = (c_hi - c_lo + CL - 1) / CL;
  const int tile0 = c_lo + (int)crank;

  if (warp == 0) {
    // ===================== producer: weight ring =====================
    if (lane == 0) {
      uint32_t st = 0, ph = 0;
      for (int slot = 0; slot < n_slots; ++slot) {
        for (int s = 0; s < DEC_LAYERS * DEC_KP * SPK; ++s) {
          umma::mbar_wait(&bars->w_empty[st], ph ^ 1u);
          umma::mbar_arrive_expect_tx(&bars->w_full[st], DEC_STAGE_B);
          if (CL == 1) {
            umma::bulk_g2s(s_w + st * DEC_STAGE_B, packed + (size_t)s * DEC_STAGE_B, DEC_STAGE_B, &bars->w_full[st]);
          } else {
            constexpr uint32_t PART = DEC_STAGE_B / CL;
            umma::bulk_g2s_multicast(s_w + st * DEC_STAGE_B + crank * PART, packed + (size_t)s * DEC_STAGE_B + crank * PART,
                                     PART, &bars->w_full[st], (uint16_t)((1u << CL) - 1u));
          }
          if (++st == NSTAGE) { st = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // "A panel kp is complete" is signalled through HARDWARE named barriers 2+kp (bar.arrive by the 512 epilogue
    // threads, bar.sync by this warp): the hand-off timeline (tools/trace_decoder.py) showed 250-500 cycles between
    // the last epilogue warp publishing a panel through an mbarrier and this thread resuming from try_wait; the
    // named barrier releases the waiting warp within tens of cycles.  MMA completion still uses tcgen05.commit ->
    // mbarrier (the only completion mechanism of the async tensor pipe).
    constexpr uint32_t idesc = MODE == MODE_BF16 ? umma::make_idesc_bf16_f32(DEC_TILE_M, DEC_H)
                                                 : umma::make_idesc_f16_f32(DEC_TILE_M, DEC_H);
    const uint32_t ah_addr = umma::smem_u32(s_ah), w_addr = umma::smem_u32(s_w);
    uint32_t st = 0, ph = 0;
    for (int slot = 0; slot < n_slots; ++slot) {
      const int tile = tile0 + slot * CL;
      if (tile >= c_hi) {
        // phantom pass (CL > 1 only): release every stage as soon as it has landed, MMA nothing
        if (lane == 0) {
          for (int s = 0; s < DEC_LAYERS * DEC_KP * SPK; ++s) {
            umma::mbar_wait(&bars->w_full[st], ph);
            for (uint32_t r = 0; r < (uint32_t)CL; ++r) umma::mbar_arrive_cluster(&bars->w_empty[st], r);
            if (++st == NSTAGE) { st = 0; ph ^= 1u; }
          }
        }
        __syncwarp();
        continue;
      }
      const bool tr = TRACE && blockIdx.x == 0 && slot < 2;
      for (int l = 0; l < DEC_LAYERS; ++l) {
        const uint32_t d = (l & 1) ? tmem_x : tmem_n;  // fc_0 -> net (fresh), fc_1 -> accumulate onto x
        for (int kp = 0; kp < DEC_KP; ++kp) {
          asm volatile("bar.sync %0, %1;" ::"r"(2 + kp), "n"(32 * DEC_EPI_WARPS + 32) : "memory");
          if (lane == 0) {
#pragma unroll
            for (int part = 0; part < SPK; ++part) {  // part 0: hi weights, part 1: lo weights
              umma::mbar_wait(&bars->w_full[st], ph);
              umma::tc_fence_after();
              if (tr && part == 0) trace[(slot * DEC_LAYERS + l) * 16 + kp] = clock64();
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t ad = umma::make_desc_k_sw128(ah_addr + kp * DEC_PANEL_A + k * 32);
                const uint64_t bd = umma::make_desc_k_sw128(w_addr + st * DEC_STAGE_B + k * 32);
                umma::mma_f16_ss(d, ad, bd, idesc, (l & 1) ? 1u : (uint32_t)((kp | k | part) != 0));
              }
              if (MODE == MODE_F16X3 && part == 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {  // a_lo . w_hi
                  const uint64_t ad = umma::make_desc_k_sw128(ah_addr + SM_AL + kp * DEC_PANEL_A + k * 32);
                  const uint64_t bd = umma::make_desc_k_sw128(w_addr + st * DEC_STAGE_B + k * 32);
                  umma::mma_f16_ss(d, ad, bd, idesc, 1u);
                }
              }
              // frees the weight slot (in every CTA of the cluster) when these MMAs retire
              if (CL == 1) umma::mma_commit(&bars->w_empty[st]);
              else umma::mma_commit_multicast(&bars->w_empty[st], (uint16_t)((1u << CL) - 1u));
              if (++st == NSTAGE) { st = 0; ph ^= 1u; }
            }
            if (kp == DEC_KP - 1) umma::mma_commit(&bars->acc_ready);
            if (tr) trace[(slot * DEC_LAYERS + l) * 16 + 4 + kp] = clock64();
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== epilogue warps (16): TMEM -> affine+ReLU -> 16-bit A panels =====================
    // TMEM is read with the 16x256b shape: a thread then owns 4 ROWS x (2 adjacent columns per 8-column block), so
    // one per-channel (a,c) fetch from shared memory serves four rows and every shared-memory store of a warp is a
    // conflict-free 128-byte wavefront (8 rows x 16 B after the 128B swizzle).  16 warps (4 per scheduler) hide the
    // TMEM / shared-memory latencies of each other; warp (q, cq) owns rows [32q, 32q+32) x columns [16cq, 16cq+16)
    // of every 64-column panel.
    const int ew = warp - 2;
    const int q = warp & 3;     // TMEM lane quarter this warp may access
    const int cq = ew >> 2;     // which 16-column quarter of every 64-column panel
    const int lr = lane >> 2;   // 0..7  row within an 8-row group  (== row & 7 : the swizzle phase)
    const int lc = lane & 3;    // 0..3  column pair within an 8-column block
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    uint32_t layer_count = 0;
    int cur_obj = -1;
    for (int slot = 0; slot < n_slots; ++slot) {
      const int tile = tile0 + slot * CL;
      if (tile >= c_hi) break;
      const int obj = tile / tiles_per_obj;
      const int t0 = (tile - obj * tiles_per_obj) * DEC_TILE_M + q * 32 + lr;  // row j of this thread: t0 + 8j
      if (obj != cur_obj) {
        // (re)load this object's x_bias + paired (a,c) table: 5888 floats, all 16 epilogue warps cooperate
        asm volatile("bar.sync 1, %0;" ::"n"(32 * DEC_EPI_WARPS) : "memory");  // everyone is done with the old table
        const float4 *src4 = reinterpret_cast<const float4 *>(aff_all + (size_t)obj * DEC_REC_FLOATS + DEC_PLAIN_FLOATS);
        float4 *dst4 = reinterpret_cast<float4 *>(s_aff);
        for (int e = tid - 64; e < DEC_AFF_FLOATS / 4; e += 32 * DEC_EPI_WARPS) dst4[e] = __ldg(src4 + e);
        asm volatile("bar.sync 1, %0;" ::"n"(32 * DEC_EPI_WARPS) : "memory");
        cur_obj = obj;
      }
      // shared-space byte addresses (explicit ld.shared / st.shared: the compiler otherwise emits generic LD/ST)
      const uint32_t aff_a = umma::smem_u32(s_aff);  // [x_bias 256][11][128 pairs]{a0,a1,c0,c1}
      const uint32_t affi_a = aff_a + DEC_H * 4 + lc * 16;                  // + layer*2048 + (col/2)*16
      const uint32_t pan0 = umma::smem_u32(s_ah) + (q * 32 + lr) * 128 + lc * 4;  // + kp*PANEL + j*1024 + swizzled chunk
      const uint32_t wp_a = umma::smem_u32(s_wp), wout_a = umma::smem_u32(s_wout);
      // ---- E0: x0 = fc_p(p) + (fc_p.bias + fc_z(z)) in fp32 -> TMEM ; h0 = relu(a0*x0 + c0) -> A panels
      {
        float px[4], py[4], pz[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int t = t0 + 8 * j;
          px[j] = py[j] = pz[j] = 0.f;
          if (t < T) {
            const float *pp = p + (size_t)obj * p_stride + (size_t)t * 3;
            px[j] = __ldg(pp); py[j] = __ldg(pp + 1); pz[j] = __ldg(pp + 2);
          }
        }
#pragma unroll 1
        for (int kp = 0; kp < DEC_KP; ++kp) {
          const int cb = kp * 64 + cq * DEC_CW;
          uint32_t v[2][8];
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int col = cb + 8 * i + 2 * lc;
            const float2 w0 = umma::lds_f2(wp_a + col * 4);
            const float2 w1 = umma::lds_f2(wp_a + (DEC_H + col) * 4);
            const float2 w2 = umma::lds_f2(wp_a + (2 * DEC_H + col) * 4);
            const float2 xb = umma::lds_f2(aff_a + col * 4);
            const float4 ac = umma::lds_f4(affi_a + (cb >> 1) * 16 + i * 64);
            const uint32_t sw = (uint32_t)(((cq * 2 + i) ^ lr) << 4);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float x0 = fmaf(px[j], w0.x, xb.x), x1 = fmaf(px[j], w0.y, xb.y);
              x0 = fmaf(py[j], w1.x, x0); x1 = fmaf(py[j], w1.y, x1);
              x0 = fmaf(pz[j], w2.x, x0); x1 = fmaf(pz[j], w2.y, x1);
              v[j >> 1][4 * i + 2 * (j & 1)] = __float_as_uint(x0);
              v[j >> 1][4 * i + 2 * (j & 1) + 1] = __float_as_uint(x1);
              store_act_pair<MODE>(pan0 + kp * DEC_PANEL_A + j * 1024 + sw, fmaf(x0, ac.x, ac.z), fmaf(x1, ac.y, ac.w));
            }
          }
          umma::tmem_st_16x256b_x2(tmem_x + lane_base + cb, v[0]);
          umma::tmem_st_16x256b_x2(tmem_x + lane_base + (16u << 16) + cb, v[1]);
          umma::tc_wait_st();
          umma::fence_proxy_async_smem();
          umma::tc_fence_before();
          asm volatile("bar.arrive %0, %1;" ::"r"(2 + kp), "n"(32 * DEC_EPI_WARPS + 32) : "memory");
        }
      }
      // ---- layers
#pragma unroll 1
      for (int l = 0; l < DEC_LAYERS; ++l, ++layer_count) {
        // (a,c) of this warp's first panel are fetched while the MMAs of the layer are still running
        const float4 pre0 = umma::lds_f4(affi_a + (l + 1) * (DEC_H * 2 * 4) + cq * (DEC_CW / 2) * 16);
        const float4 pre1 = umma::lds_f4(affi_a + (l + 1) * (DEC_H * 2 * 4) + cq * (DEC_CW / 2) * 16 + 64);
        umma::mbar_wait(&bars->acc_ready, layer_count & 1u);
        umma::tc_fence_after();
        const bool tr = TRACE && blockIdx.x == 0 && slot < 2 && warp == 2 && lane == 0;
        if (tr) trace[(slot * DEC_LAYERS + l) * 16 + 8] = clock64();
        const uint32_t src = ((l & 1) ? tmem_x : tmem_n) + lane_base + cq * DEC_CW;
        const uint32_t al = affi_a + (l + 1) * (DEC_H * 2 * 4) + cq * (DEC_CW / 2) * 16;  // + kp*512 + i*64
        // software pipeline over the four panels: the TMEM load of panel kp+1 is in flight while panel kp is
        // converted, stored and published
        uint32_t v[2][2][8];
        umma::tmem_ld_16x256b_x2(src, v[0][0]);
        umma::tmem_ld_16x256b_x2(src + (16u << 16), v[0][1]);
        if (l < DEC_LAYERS - 1) {
#pragma unroll
          for (int kp = 0; kp < DEC_KP; ++kp) {
            const float4 ac0 = kp ? umma::lds_f4(al + kp * 512) : pre0, ac1 = kp ? umma::lds_f4(al + kp * 512 + 64) : pre1;
            umma::tc_wait_ld();
            if (tr && kp == 0) trace[(slot * DEC_LAYERS + l) * 16 + 13] = clock64();
            if (kp + 1 < DEC_KP) {
              umma::tmem_ld_16x256b_x2(src + (kp + 1) * 64, v[(kp + 1) & 1][0]);
              umma::tmem_ld_16x256b_x2(src + (16u << 16) + (kp + 1) * 64, v[(kp + 1) & 1][1]);
            }
            const uint32_t pan = pan0 + kp * DEC_PANEL_A;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const float4 ac = i ? ac1 : ac0;
              const uint32_t sw = (uint32_t)(((cq * 2 + i) ^ lr) << 4);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float x0 = __uint_as_float(v[kp & 1][j >> 1][4 * i + 2 * (j & 1)]);
                const float x1 = __uint_as_float(v[kp & 1][j >> 1][4 * i + 2 * (j & 1) + 1]);
                store_act_pair<MODE>(pan + j * 1024 + sw, fmaf(x0, ac.x, ac.z), fmaf(x1, ac.y, ac.w));
              }
            }
            if (tr && kp == 0) trace[(slot * DEC_LAYERS + l) * 16 + 14] = clock64();
            umma::fence_proxy_async_smem();
            if (tr && kp == 0) trace[(slot * DEC_LAYERS + l) * 16 + 15] = clock64();
            umma::tc_fence_before();
            asm volatile("bar.arrive %0, %1;" ::"r"(2 + kp), "n"(32 * DEC_EPI_WARPS + 32) : "memory");
            if (tr) trace[(slot * DEC_LAYERS + l) * 16 + 9 + kp] = clock64();
          }
        } else {
          // final: logits = fc_out(relu(cbn(x)))  -- fp32 dot product over the 256 channels
          float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int kp = 0; kp < DEC_KP; ++kp) {
            const float4 acs[2] = {umma::lds_f4(al + kp * 512), umma::lds_f4(al + kp * 512 + 64)};
            const float2 wos[2] = {umma::lds_f2(wout_a + (kp * 64 + cq * DEC_CW + 2 * lc) * 4),
                                   umma::lds_f2(wout_a + (kp * 64 + cq * DEC_CW + 8 + 2 * lc) * 4)};
            umma::tc_wait_ld();
            if (kp + 1 < DEC_KP) {
              umma::tmem_ld_16x256b_x2(src + (kp + 1) * 64, v[(kp + 1) & 1][0]);
              umma::tmem_ld_16x256b_x2(src + (16u << 16) + (kp + 1) * 64, v[(kp + 1) & 1][1]);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float x0 = __uint_as_float(v[kp & 1][j >> 1][4 * i + 2 * (j & 1)]);
                const float x1 = __uint_as_float(v[kp & 1][j >> 1][4 * i + 2 * (j & 1) + 1]);
                part[j] = fmaf(fmaxf(fmaf(x0, acs[i].x, acs[i].z), 0.f), wos[i].x, part[j]);
                part[j] = fmaf(fmaxf(fmaf(x1, acs[i].y, acs[i].w), 0.f), wos[i].y, part[j]);
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            part[j] += __shfl_xor_sync(0xffffffffu, part[j], 1);
            part[j] += __shfl_xor_sync(0xffffffffu, part[j], 2);
            if (lc == 0) s_out[cq * DEC_TILE_M + q * 32 + lr + 8 * j] = part[j];
          }
          umma::tc_fence_before();
          asm volatile("bar.sync 1, %0;" ::"n"(32 * DEC_EPI_WARPS) : "memory");  // the epilogue warps only
          if (cq == 0 && lc == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int t = t0 + 8 * j, rr = q * 32 + lr + 8 * j;
              if (t < T)
                logits[(size_t)obj * T + t] =
                    ((s_out[rr] + s_out[DEC_TILE_M + rr]) + (s_out[2 * DEC_TILE_M + rr] + s_out[3 * DEC_TILE_M + rr])) + fc_out_b;
            }
          }
        }
      }
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (CL > 1) umma::cluster_sync();  // no CTA exits while a peer may still multicast into it / arrive on its barriers
  if (warp == 1) {
    umma::tc_fence_after();
    umma::tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// weight packing: fc_w (10,256,256) f32 [layer][out n][in k]  ->  bf16, per (layer, k-panel) a 32-KB image of the
// [256 rows n][64 k] K-major SWIZZLE_128B operand, stages in consumption order.
// MODE_F16X3: two stages per (layer, k-panel): hi = fp16(w), lo = fp16(w - hi).
template <int MODE>
__global__ void onet_pack_kernel(const float *__restrict__ fc_w, uint8_t *__restrict__ packed) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;  // one thread per 8 consecutive k (16 bytes)
  const int total = DEC_LAYERS * DEC_H * (DEC_H / 8);
  if (e >= total) return;
  const int l = e / (DEC_H * 32), rem = e % (DEC_H * 32);
  const int n = rem / 32, kc = rem % 32;  // kc: 16-byte chunk index along k (0..31)
  const int kp = kc / 8, cin = kc % 8;
  const float *src = fc_w + ((size_t)l * DEC_H + n) * DEC_H + kc * 8;
  constexpr int SPK = MODE == MODE_F16X3 ? 2 : 1;
  uint32_t w[4], wl[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (MODE == MODE_BF16) {
      w[i] = umma::pack_bf16x2(src[2 * i], src[2 * i + 1]);
    } else {
      w[i] = umma::pack_f16x2(src[2 * i], src[2 * i + 1]);
      const float2 hf = umma::unpack_f16x2(w[i]);
      wl[i] = umma::pack_f16x2(src[2 * i] - hf.x, src[2 * i + 1] - hf.y);
    }
  }
  uint8_t *dst = packed + (size_t)((l * DEC_KP + kp) * SPK) * DEC_STAGE_B + n * 128 + ((cin ^ (n & 7)) << 4);
  *reinterpret_cast<uint4 *>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
  if (MODE == MODE_F16X3) *reinterpret_cast<uint4 *>(dst + DEC_STAGE_B) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
}

// ------------------------------------------------------------------------------------------------
// conditional-BN tables.  grid (11 layers, ceil(B/8)), 256 threads; each warp produces 32 channels for 8 objects.
constexpr int CBN_OBJ = 8;
__global__ void __launch_bounds__(256)
onet_cbn_tables_kernel(const float *__restrict__ c, int B, int c_dim, const float *__restrict__ gamma_w,
                       const float *__restrict__ gamma_b, const float *__restrict__ beta_w,
                       const float *__restrict__ beta_b, const float *__restrict__ run_mean,
                       const float *__restrict__ run_var, float eps, const float *__restrict__ fc_bias,
                       const float *__restrict__ x_bias, float *__restrict__ aff) {
  extern __shared__ float s_c[];  // [CBN_OBJ][c_dim]
  const int l = blockIdx.x, b0 = blockIdx.y * CBN_OBJ;
  const int nb = min(CBN_OBJ, B - b0);
  for (int e = threadIdx.x; e < CBN_OBJ * c_dim; e += blockDim.x) {
    const int o = e / c_dim, k = e % c_dim;
    s_c[e] = o < nb ? __ldg(c + (size_t)(b0 + o) * c_dim + k) : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int ci = 0; ci < 32; ++ci) {
    const int ch = warp * 32 + ci;
    const float *gw = gamma_w + ((size_t)l * DEC_H + ch) * c_dim;
    const float *bw = beta_w + ((size_t)l * DEC_H + ch) * c_dim;
    float g[CBN_OBJ], bt[CBN_OBJ];
#pragma unroll
    for (int o = 0; o < CBN_OBJ; ++o) { g[o] = 0.f; bt[o] = 0.f; }
    for (int k = lane; k < c_dim; k += 32) {
      const float wg = __ldg(gw + k), wb = __ldg(bw + k);
#pragma unroll
      for (int o = 0; o < CBN_OBJ; ++o) {
        const float cv = s_c[o * c_dim + k];
        g[o] = fmaf(wg, cv, g[o]);
        bt[o] = fmaf(wb, cv, bt[o]);
      }
    }
#pragma unroll
    for (int o = 0; o < CBN_OBJ; ++o) {
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        g[o] += __shfl_xor_sync(0xffffffffu, g[o], off);
        bt[o] += __shfl_xor_sync(0xffffffffu, bt[o], off);
      }
    }
    if (lane < nb) {
      float gam = 0.f, bet = 0.f;
#pragma unroll
      for (int o = 0; o < CBN_OBJ; ++o)
        if (o == lane) { gam = g[o]; bet = bt[o]; }
      gam += __ldg(gamma_b + l * DEC_H + ch);
      bet += __ldg(beta_b + l * DEC_H + ch);
      // bias still pending on the accumulator this CBN reads:
      //   odd l  (bn_1 of block (l-1)/2): the bias of fc_0 = fc layer l-1
      //   even l (bn_0 of block l/2, or the final bn): sum of the fc_1 biases of all previous blocks
      float pend = 0.f;
      if (l & 1) {
        pend = __ldg(fc_bias + (l - 1) * DEC_H + ch);
      } else {
        for (int tblk = 0; tblk < l / 2; ++tblk) pend += __ldg(fc_bias + (2 * tblk + 1) * DEC_H + ch);
      }
      const float a = gam * rsqrtf(__ldg(run_var + l * DEC_H + ch) + eps);
      const float cc = fmaf(a, pend - __ldg(run_mean + l * DEC_H + ch), bet);
      float *rec = aff + (size_t)(b0 + lane) * DEC_REC_FLOATS;
      rec[l * 2 * DEC_H + ch] = a;
      rec[l * 2 * DEC_H + DEC_H + ch] = cc;
      if (l == 0) rec[DEC_PLAIN_FLOATS + ch] = __ldg(x_bias + (size_t)(b0 + lane) * DEC_H + ch);
      float *pr = rec + DEC_PLAIN_FLOATS + DEC_H + l * 2 * DEC_H + (ch >> 1) * 4 + (ch & 1);
      pr[0] = a;
      pr[2] = cc;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// fp32 exact path pieces
__global__ void dec_fcp_kernel(const float *__restrict__ p, long long p_stride, int T, const float *__restrict__ fc_p_w,
                               const float *__restrict__ aff, int b0, float *__restrict__ x) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int o = blockIdx.y, bl = blockIdx.z;  // bl: object within chunk
  if (t >= T) return;
  const float *pp = p + (size_t)(b0 + bl) * p_stride + (size_t)t * 3;
  const float xb = __ldg(aff + (size_t)(b0 + bl) * DEC_REC_FLOATS + DEC_PLAIN_FLOATS + o);
  float acc = fmaf(__ldg(pp), __ldg(fc_p_w + o * 3), xb);
  acc = fmaf(__ldg(pp + 1), __ldg(fc_p_w + o * 3 + 1), acc);
  acc = fmaf(__ldg(pp + 2), __ldg(fc_p_w + o * 3 + 2), acc);
  x[((size_t)bl * DEC_H + o) * T + t] = acc;
}

__global__ void dec_out_kernel(const float *__restrict__ x, int T, const float *__restrict__ aff, int b0,
                               const float *__restrict__ fc_out_w, float fc_out_b, float *__restrict__ logits) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int bl = blockIdx.y;
  if (t >= T) return;
  const float *rec = aff + (size_t)(b0 + bl) * DEC_REC_FLOATS + 10 * 2 * DEC_H;
  float acc = 0.f;
  for (int k = 0; k < DEC_H; ++k) {
    const float h = fmaxf(fmaf(__ldg(x + ((size_t)bl * DEC_H + k) * T + t), __ldg(rec + k), __ldg(rec + DEC_H + k)), 0.f);
    acc = fmaf(h, __ldg(fc_out_w + k), acc);
  }
  logits[(size_t)(b0 + bl) * T + t] = acc + fc_out_b;
}

// ------------------------------------------------------------------------------------------------
// tcgen05 self-test: D (128x256 f32) = bf16(A (128x64)) . bf16(B (256x64))^T through exactly the descriptor,
// swizzle, MMA, commit and TMEM-load helpers the decoder uses.
__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const float *__restrict__ A, const float *__restrict__ Bm, float *__restrict__ D) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *sa = smem, *sb = smem + DEC_PANEL_A;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { umma::mbar_init(&bar, 1); umma::fence_barrier_init(); }
  if (warp == 0) umma::tmem_alloc(&tmem_ptr, 256);
  for (int e = tid; e < 128 * 32; e += 128) {
    const int row = e / 32, k2 = (e % 32) * 2;
    *reinterpret_cast<uint32_t *>(sa + umma::sw128_offset(row, k2)) = umma::pack_bf16x2(A[row * 64 + k2], A[row * 64 + k2 + 1]);
  }
  for (int e = tid; e < 256 * 32; e += 128) {
    const int row = e / 32, k2 = (e % 32) * 2;
    *reinterpret_cast<uint32_t *>(sb + umma::sw128_offset(row, k2)) = umma::pack_bf16x2(Bm[row * 64 + k2], Bm[row * 64 + k2 + 1]);
  }
  umma::fence_proxy_async_smem();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tb = tmem_ptr;
  if (tid == 0) {
    constexpr uint32_t idesc = umma::make_idesc_bf16_f32(128, 256);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma::mma_bf16_ss(tb, umma::make_desc_k_sw128(umma::smem_u32(sa) + k * 32),
                        umma::make_desc_k_sw128(umma::smem_u32(sb) + k * 32), idesc, k != 0);
    umma::mma_commit(&bar);
  }
  umma::mbar_wait(&bar, 0);
  umma::tc_fence_after();
  for (int c0 = 0; c0 < 256; c0 += 32) {
    uint32_t v[32];
    umma::tmem_ld32(tb + ((uint32_t)(warp * 32) << 16) + c0, v);
    umma::tc_wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * 256 + c0 + j] = __uint_as_float(v[j]);
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) { umma::tc_fence_after(); umma::tmem_dealloc(tb, 256); }
}

// TS-mode self-test: same product, but A is first written into tensor memory (packed bf16 pairs) with
// tcgen05.st.16x128b by threads that own (row, column-pair) positions in the 16x256b accumulator layout -- exactly what
// the decoder's epilogue does -- and the MMA reads A from TMEM.
__global__ void __launch_bounds__(128, 1)
umma_selftest_ts_kernel(const float *__restrict__ A, const float *__restrict__ Bm, float *__restrict__ D) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *sb = smem;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lc = lane & 3;
  if (tid == 0) { umma::mbar_init(&bar, 1); umma::fence_barrier_init(); }
  if (warp == 0) umma::tmem_alloc(&tmem_ptr, 512);
  for (int e = tid; e < 256 * 32; e += 128) {
    const int row = e / 32, k2 = (e % 32) * 2;
    *reinterpret_cast<uint32_t *>(sb + umma::sw128_offset(row, k2)) = umma::pack_bf16x2(Bm[row * 64 + k2], Bm[row * 64 + k2 + 1]);
  }
  umma::fence_proxy_async_smem();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tb = tmem_ptr;
  const uint32_t ta = tb + 256;  // A lives in columns [256, 288): 64 k = 32 packed columns
  // warp w owns lanes [32w, 32w+32); K = 64 -> four 16-column groups g (8 packed columns each)
  for (int g = 0; g < 4; ++g) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t pk[2][2];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int row = warp * 32 + half * 16 + lr + 8 * jj;
          const int k = g * 16 + 8 * i + 2 * lc;
          pk[i][jj] = umma::pack_bf16x2(A[row * 64 + k], A[row * 64 + k + 1]);
        }
      umma::tmem_st_16x128b_x2(ta + ((uint32_t)(warp * 32 + half * 16) << 16) + g * 8, pk[0][0], pk[0][1], pk[1][0], pk[1][1]);
    }
  }
  umma::tc_wait_st();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  if (tid == 0) {
    constexpr uint32_t idesc = umma::make_idesc_bf16_f32(128, 256);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma::mma_bf16_ts(tb, ta + k * 8, umma::make_desc_k_sw128(umma::smem_u32(sb) + k * 32), idesc, k != 0);
    umma::mma_commit(&bar);
  }
  umma::mbar_wait(&bar, 0);
  umma::tc_fence_after();
  for (int c0 = 0; c0 < 256; c0 += 32) {
    uint32_t v[32];
    umma::tmem_ld32(tb + ((uint32_t)(warp * 32) << 16) + c0, v);
    umma::tc_wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * 256 + c0 + j] = __uint_as_float(v[j]);
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) { umma::tc_fence_after(); umma::tmem_dealloc(tb, 512); }
}

}  // namespace rfd

using namespace rfd;

extern "C" int rfd_umma_selftest_ts(const float *A, const float *Bm, float *D, void *stream) {
  if (!A || !Bm || !D) return RFD_ERR_INVALID_ARGUMENT;
  const int smem = DEC_H * 128 + 1024;
  RFD_CHECK_CUDA(cudaFuncSetAttribute(umma_selftest_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem),
                 "selftest_ts attr");
  umma_selftest_ts_kernel<<<1, 128, smem, as_stream(stream)>>>(A, Bm, D);
  RFD_CHECK_LAUNCH("umma_selftest_ts_kernel");
  return RFD_OK;
}

static bool valid_mode(int mode) { return mode == MODE_BF16 || mode == MODE_F16 || mode == MODE_F16X3; }

extern "C" size_t rfd_onet_packed_bytes(int mode) {
  if (!valid_mode(mode)) return 0;
  return (size_t)DEC_LAYERS * DEC_KP * DEC_STAGE_B * (mode == MODE_F16X3 ? 2 : 1);
}

extern "C" size_t rfd_onet_aff_floats(void) { return DEC_REC_FLOATS; }

extern "C" int rfd_onet_pack_weights(const float *fc_w, int mode, void *packed, void *stream) {
  if (!fc_w || !packed || !valid_mode(mode)) return RFD_ERR_INVALID_ARGUMENT;
  const int total = DEC_LAYERS * DEC_H * (DEC_H / 8);
  const int grid = h_ceil_div(total, 256);
  uint8_t *dst = reinterpret_cast<uint8_t *>(packed);
  if (mode == MODE_BF16) onet_pack_kernel<MODE_BF16><<<grid, 256, 0, as_stream(stream)>>>(fc_w, dst);
  else if (mode == MODE_F16) onet_pack_kernel<MODE_F16><<<grid, 256, 0, as_stream(stream)>>>(fc_w, dst);
  else onet_pack_kernel<MODE_F16X3><<<grid, 256, 0, as_stream(stream)>>>(fc_w, dst);
  RFD_CHECK_LAUNCH("onet_pack_kernel");
  return RFD_OK;
}

extern "C" int rfd_onet_cbn_tables(const float *c, int B, int c_dim, const float *gamma_w, const float *gamma_b,
                                   const float *beta_w, const float *beta_b, const float *run_mean,
                                   const float *run_var, float eps, const float *fc_bias, const float *x_bias,
                                   float *aff, void *stream) {
  if (B < 0 || c_dim < 1) return RFD_ERR_INVALID_ARGUMENT;
  if (B == 0) return RFD_OK;
  if (!c || !gamma_w || !gamma_b || !beta_w || !beta_b || !run_mean || !run_var || !fc_bias || !x_bias || !aff)
    return RFD_ERR_INVALID_ARGUMENT;
  const size_t smem = sizeof(float) * CBN_OBJ * (size_t)c_dim;
  if (smem > 96 * 1024 || h_ceil_div(B, CBN_OBJ) > 65535) return RFD_ERR_UNSUPPORTED_SIZE;
  if (smem > 48 * 1024)
    RFD_CHECK_CUDA(cudaFuncSetAttribute(onet_cbn_tables_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                   "cbn attr");
  dim3 grid(DEC_CBN, h_ceil_div(B, CBN_OBJ));
  onet_cbn_tables_kernel<<<grid, 256, smem, as_stream(stream)>>>(c, B, c_dim, gamma_w, gamma_b, beta_w, beta_b, run_mean,
                                                                  run_var, eps, fc_bias, x_bias, aff);
  RFD_CHECK_LAUNCH("onet_cbn_tables_kernel");
  return RFD_OK;
}

// weight-sharing cluster size of the tensor-core decoder (1 = independent CTAs, 2 = CTA pairs with multicast weight
// stages).  Process-wide tuning knob; the default comes from RFD_ONET_CLUSTER (else DEC_DEFAULT_CLUSTER).
constexpr int DEC_DEFAULT_CLUSTER = 2;
static std::atomic<int> g_decode_cluster{0};

static int decode_cluster() {
  int c = g_decode_cluster.load(std::memory_order_relaxed);
  if (c == 0) {
    const char *e = getenv("RFD_ONET_CLUSTER");
    c = (e && (e[0] == '1' || e[0] == '2') && e[1] == 0) ? e[0] - '0' : DEC_DEFAULT_CLUSTER;
    g_decode_cluster.store(c, std::memory_order_relaxed);
  }
  return c;
}

extern "C" int rfd_onet_decode_set_cluster(int cluster) {
  if (cluster != 1 && cluster != 2) return RFD_ERR_INVALID_ARGUMENT;
  g_decode_cluster.store(cluster, std::memory_order_relaxed);
  return RFD_OK;
}

template <int MODE, int CL, bool TRACE>
static int launch_decode_t(const float *p, long long p_stride, int T, const float *fc_p_w, const uint8_t *packed,
                           const float *aff, const float *fc_out_w, float fc_out_b, float *logits, int num_tiles,
                           int tiles_per_obj, unsigned long long *trace, int sms, cudaStream_t st) {
  auto kern = onet_decode_kernel<MODE, CL, TRACE>;
  RFD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, DEC_SMEM_BYTES), "decode attr");
  int grid = num_tiles < sms ? num_tiles : sms;
  grid = (grid / CL) * CL;  // whole clusters (CL > 1 callers guarantee num_tiles >= CL)
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(DEC_THREADS);
  cfg.dynamicSmemBytes = DEC_SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  RFD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, p, p_stride, T, fc_p_w, packed, aff, fc_out_w, fc_out_b, logits, num_tiles,
                                    tiles_per_obj, trace),
                 "onet_decode_kernel launch");
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return RFD_OK;
}

static int launch_decode(const float *p, long long p_batch_stride, int B, int T, const float *fc_p_w, const void *packed,
                         int mode, const float *aff, const float *fc_out_w, float fc_out_b, float *logits,
                         unsigned long long *trace, void *stream) {
  if (B < 0 || T < 0 || p_batch_stride < 0) return RFD_ERR_INVALID_ARGUMENT;
  if (!valid_mode(mode)) return RFD_ERR_INVALID_ARGUMENT;
  if (B == 0 || T == 0) return RFD_OK;
  if (!p || !fc_p_w || !packed || !aff || !fc_out_w || !logits) return RFD_ERR_INVALID_ARGUMENT;
  const long long tiles_per_obj = (T + DEC_TILE_M - 1) / DEC_TILE_M;
  const long long num_tiles = tiles_per_obj * B;
  if (num_tiles > 0x7fffffffLL) return RFD_ERR_UNSUPPORTED_SIZE;
  int dev = 0, sms = 148;
  RFD_CHECK_CUDA(cudaGetDevice(&dev), "decode getdevice");
  RFD_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "decode sms");
  const uint8_t *pk = reinterpret_cast<const uint8_t *>(packed);
  cudaStream_t st = as_stream(stream);
  const int nt = (int)num_tiles, tpo = (int)tiles_per_obj;
  const int cl = (num_tiles >= 2 && !trace) ? decode_cluster() : 1;
#define RFD_DEC_LAUNCH(M, C, TR) \
  return launch_decode_t<M, C, TR>(p, p_batch_stride, T, fc_p_w, pk, aff, fc_out_w, fc_out_b, logits, nt, tpo, trace, sms, st)
  if (trace) {
    if (mode != MODE_F16) return RFD_ERR_INVALID_ARGUMENT;  // the instrumented instantiation exists for the default mode
    RFD_DEC_LAUNCH(MODE_F16, 1, true);
  }
  if (mode == MODE_BF16) { if (cl == 2) RFD_DEC_LAUNCH(MODE_BF16, 2, false); RFD_DEC_LAUNCH(MODE_BF16, 1, false); }
  if (mode == MODE_F16) { if (cl == 2) RFD_DEC_LAUNCH(MODE_F16, 2, false); RFD_DEC_LAUNCH(MODE_F16, 1, false); }
  if (cl == 2) RFD_DEC_LAUNCH(MODE_F16X3, 2, false);
  RFD_DEC_LAUNCH(MODE_F16X3, 1, false);
#undef RFD_DEC_LAUNCH
}

extern "C" int rfd_onet_decode(const float *p, long long p_batch_stride, int B, int T, const float *fc_p_w,
                               const void *packed, int mode, const float *aff, const float *fc_out_w,
                               float fc_out_b, float *logits, void *stream) {
  return launch_decode(p, p_batch_stride, B, T, fc_p_w, packed, mode, aff, fc_out_w, fc_out_b, logits, nullptr, stream);
}

// diagnostics: the same decode with an instrumented kernel (RFD_ONET_MODE_F16, cluster 1) whose CTA 0 records the hand-off
// timeline of its first two tiles into trace (device, 2*10*16 u64).  The trace buffer is an argument, not process state.
extern "C" int rfd_onet_decode_traced(const float *p, long long p_batch_stride, int B, int T, const float *fc_p_w,
                                      const void *packed, int mode, const float *aff, const float *fc_out_w,
                                      float fc_out_b, float *logits, unsigned long long *trace, void *stream) {
  if (!trace) return RFD_ERR_INVALID_ARGUMENT;
  return launch_decode(p, p_batch_stride, B, T, fc_p_w, packed, mode, aff, fc_out_w, fc_out_b, logits, trace, stream);
}

extern "C" int rfd_onet_decode_f32(const float *p, long long p_batch_stride, int B, int T, const float *fc_p_w,
                                   const float *fc_w, const float *aff, const float *fc_out_w, float fc_out_b,
                                   float *logits, float *workspace, size_t workspace_bytes, void *stream) {
  if (B < 0 || T < 0 || p_batch_stride < 0) return RFD_ERR_INVALID_ARGUMENT;
  if (B == 0 || T == 0) return RFD_OK;
  if (!p || !fc_p_w || !fc_w || !aff || !fc_out_w || !logits || !workspace) return RFD_ERR_INVALID_ARGUMENT;
  const size_t per_obj = (size_t)2 * DEC_H * T * sizeof(float);
  long long bc = (long long)(workspace_bytes / per_obj);
  if (bc < 1) return RFD_ERR_INVALID_ARGUMENT;
  if (bc > B) bc = B;
  if (bc > 65535) bc = 65535;
  cudaStream_t st = as_stream(stream);
  // unit scale / zero shift for the bias-free layers: one small constant buffer per device, created on first use
  static float *ones_zeros_dev[64] = {nullptr};
  static std::mutex ones_zeros_mu;
  int cur_dev = 0;
  RFD_CHECK_CUDA(cudaGetDevice(&cur_dev), "decode_f32 getdevice");
  if (cur_dev < 0 || cur_dev >= 64) return RFD_ERR_UNSUPPORTED_SIZE;
  float *ones_zeros = nullptr;
  {
    std::lock_guard<std::mutex> lk(ones_zeros_mu);
    if (!ones_zeros_dev[cur_dev]) {
      float h[2 * DEC_H];
      for (int i = 0; i < DEC_H; ++i) { h[i] = 1.f; h[DEC_H + i] = 0.f; }
      float *d = nullptr;
      RFD_CHECK_CUDA(cudaMalloc(&d, sizeof(h)), "decode_f32 malloc");
      RFD_CHECK_CUDA(cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice), "decode_f32 memcpy");
      ones_zeros_dev[cur_dev] = d;
    }
    ones_zeros = ones_zeros_dev[cur_dev];
  }
  for (int b0 = 0; b0 < B; b0 += (int)bc) {
    const int nb = (int)((B - b0) < bc ? (B - b0) : bc);
    float *x = workspace, *net = workspace + (size_t)bc * DEC_H * T;
    dec_fcp_kernel<<<dim3(h_ceil_div(T, 256), DEC_H, nb), 256, 0, st>>>(p, p_batch_stride, T, fc_p_w, aff, b0, x);
    RFD_CHECK_LAUNCH("dec_fcp_kernel");
    const float *affb = aff + (size_t)b0 * DEC_REC_FLOATS;
    for (int i = 0; i < 5; ++i) {
      // net = W_{2i} . relu(a_{2i} x + c_{2i})
      int rc = launch_pointwise_f32(x, fc_w + (size_t)(2 * i) * DEC_H * DEC_H, ones_zeros, ones_zeros + DEC_H, nullptr,
                                    affb + (2 * i) * 2 * DEC_H, affb + (2 * i) * 2 * DEC_H + DEC_H, DEC_REC_FLOATS, 0, 1,
                                    nb, DEC_H, DEC_H, T, net, st);
      if (rc != RFD_OK) return rc;
      // x = x + W_{2i+1} . relu(a_{2i+1} net + c_{2i+1})     (in place: each element read then written by one thread)
      rc = launch_pointwise_f32(net, fc_w + (size_t)(2 * i + 1) * DEC_H * DEC_H, ones_zeros, ones_zeros + DEC_H, x,
                                affb + (2 * i + 1) * 2 * DEC_H, affb + (2 * i + 1) * 2 * DEC_H + DEC_H, DEC_REC_FLOATS,
                                0, 1, nb, DEC_H, DEC_H, T, x, st);
      if (rc != RFD_OK) return rc;
    }
    dec_out_kernel<<<dim3(h_ceil_div(T, 256), nb), 256, 0, st>>>(x, T, aff, b0, fc_out_w, fc_out_b, logits);
    RFD_CHECK_LAUNCH("dec_out_kernel");
  }
  return RFD_OK;
}

extern "C" int rfd_umma_selftest(const float *A, const float *Bm, float *D, void *stream) {
  if (!A || !Bm || !D) return RFD_ERR_INVALID_ARGUMENT;
  const int smem = DEC_PANEL_A + DEC_STAGE_B + 1024;
  RFD_CHECK_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem),
                 "selftest attr");
  umma_selftest_kernel<<<1, 128, smem, as_stream(stream)>>>(A, Bm, D);
  RFD_CHECK_LAUNCH("umma_selftest_kernel");
  return RFD_OK;
}
