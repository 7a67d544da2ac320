// fps.cu -- furthest point sampling for sm_100a.
//
// Replaces furthest_point_sampling_kernel (reference _ext-src/src/sampling_gpu.cu:69-173, launcher :175-229,
// wrapper sampling.cpp:66-87).  The reference runs ONE 512-thread CTA per scene and, per round, streams all N
// points plus the (B,N) `temp` scratch through that single SM, with a 9-stage shared-memory tree behind 10
// __syncthreads.  Here a scene is owned by a thread-block CLUSTER (up to 16 CTAs = 16 SMs):
//   * every thread keeps its points' xyz and running min-distance in REGISTERS for the whole kernel
//     (no `temp` tensor exists; HBM traffic is the compulsory 12*N bytes in + 4*m bytes out);
//   * per round: register update -> warp arg-max with two redux.sync -> one __syncthreads -> CTA arg-max ->
//     the CTA winner (key + xyz) is pushed to every peer CTA's shared memory over DSMEM and signalled with a
//     the stores' own mbarrier complete_tx (st.async); peers wait on their local mbarrier.
//     No cluster-wide barrier and no global-memory round trip sits on the serial chain.
//
// Bit-exactness with the reference:
//   * arithmetic order of the sm_100 build of the reference (cuobjdump): mag = fma(z,z,fma(x,x,y*y)),
//     skip iff (double)mag <= 1e-3; d = fma(dz,dz,fma(dx,dx,dy*dy)) with dx = p - p_old; temp = fminf(d,temp);
//   * arg-max tie-break.  Reference thread t (block size bs = opt_n_threads(N), cuda_utils.h:15-19) scans
//     k = t, t+bs, ... with a strict '>' (first max wins, :108-109); the tree (:115-168, __update :59-65) keeps
//     the LOWER slot on ties at each stage, stage strides bs/2 ... 1.  The overall winner among equal values
//     is therefore the candidate with the smallest  rank(k) = bitrev_{log2 bs}(k mod bs) * ceil(N/bs) + k div bs.
//     We reduce the 64-bit key (float_bits(value) : ~rank) with max, which is order independent.
//     A thread here owns k = g, g+T, g+2T, ... (T = threads per cluster, a multiple of bs), i.e. one slot in
//     ascending k -- so the in-register strict '>' scan is already rank-ordered;
//   * "no candidate" (all points skipped) reproduces the reference result old = 0.
#include <atomic>

#include "common.cuh"
#include <cmath>
#include <cstdio>
#include <cstdlib>

namespace rfd {

constexpr int FPS_MAX_WARPS = 32;
constexpr int FPS_MAX_CS = 16;
constexpr int FPS_MAX_PPT = 24;    // 512-thread CTAs (<= 128 registers/thread)
constexpr int FPS_MAX_PPT_1K = 8;  // 1024-thread CTAs (<= 64 registers/thread)

struct __align__(16) FpsRec {
  uint32_t hi, lo;  // key: value bits, ~rank  (0,0 = no candidate)
  int k, pad0;
  float x, y, z, pad1;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t addr, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
// 16-byte asynchronous store into a peer CTA's shared memory that, on completion, performs complete_tx(16) on
// the peer's mbarrier: data and signal travel together, no release fence (an `mbarrier.arrive.release.cluster`
// after plain st.shared::cluster stores costs an ERRBAR/MEMBAR on the serial chain of every round).
__device__ __forceinline__ void st_async_v4(uint32_t cluster_addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d,
                                            uint32_t cluster_mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                   cluster_addr),
               "r"(a), "r"(b), "r"(c), "r"(d), "r"(cluster_mbar)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_local(uint32_t addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
// Wait on the LOCAL mbarrier.  Deliberately the default (.acquire.cta) form, as in CUTLASS's ClusterBarrier::wait:
// the peers' records are written straight into THIS CTA's shared memory (st.shared::cluster) before their
// release.cluster arrive, so nothing cached needs invalidating -- whereas `.acquire.cluster` makes ptxas emit a
// CCTL.IVALL (L1 invalidate-all) after every wait, which was 45% of all stall samples of this kernel (profiles/).
__device__ __forceinline__ void mbar_wait_acquire_cluster(uint32_t addr, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(addr),
      "r"(parity)
      : "memory");
}

// warp-level arg-max over up to 32 records in shared memory; returns the winning record in every lane
__device__ __forceinline__ FpsRec fps_pick(const FpsRec *recs, int n, int lane) {
  uint32_t hi = 0, lo = 0;
  if (lane < n) {
    const uint2 h = *reinterpret_cast<const uint2 *>(&recs[lane]);
    hi = h.x;
    lo = h.y;
  }
  const uint32_t hmax = __reduce_max_sync(0xffffffffu, hi);
  const uint32_t lmax = __reduce_max_sync(0xffffffffu, hi == hmax ? lo : 0u);
  const uint32_t win = __ballot_sync(0xffffffffu, lane < n && hi == hmax && lo == lmax);
  const int src = __ffs(win) - 1;
  const uint4 a = *reinterpret_cast<const uint4 *>(&recs[src]);
  const float4 b = *(reinterpret_cast<const float4 *>(&recs[src]) + 1);
  FpsRec r;
  r.hi = a.x; r.lo = a.y; r.k = (int)a.z; r.pad0 = 0;
  r.x = b.x; r.y = b.y; r.z = b.z; r.pad1 = 0.f;
  return r;
}

template <int PPT, int FPS_THREADS>
__global__ void __launch_bounds__(FPS_THREADS, 1)
fps_kernel(const float *__restrict__ xyz_all, int N, int m, int bs_log2, int Q, int CS, int *__restrict__ idx_all,
           float *__restrict__ new_xyz_all, const int *__restrict__ prefix_flag) {
  extern __shared__ float4 s_pts[];  // [PPT][FPS_THREADS] this CTA's points (winner looks its xyz up here)
  constexpr int FPS_WARPS = FPS_THREADS / 32;
  __shared__ FpsRec s_warp[2][FPS_WARPS];
  __shared__ FpsRec s_cta[2][FPS_MAX_CS];
  __shared__ __align__(8) uint64_t s_mbar[2];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = (CS > 1) ? cluster_ctarank() : 0u;
  const int scene = blockIdx.x / CS;
  const float *__restrict__ xyz = xyz_all + (size_t)scene * N * 3;
  int *__restrict__ idx_out = idx_all + (size_t)scene * m;
  const int T = CS * FPS_THREADS;
  const int g = (int)rank * FPS_THREADS + tid;

  // rfd_fps_prefix_check proved that sampling this scene returns 0, 1, ..., m-1 (the input is itself in FPS order):
  // emit the identity and skip the m-1 serial rounds.  Uniform over the scene's cluster, before any cluster traffic.
  if (prefix_flag && __ldg(prefix_flag + scene) != 0) {
    for (int j = g; j < m; j += T) {
      idx_out[j] = j;
      if (new_xyz_all) {
        float *o = new_xyz_all + ((size_t)scene * m + j) * 3;
        o[0] = __ldg(xyz + (size_t)j * 3); o[1] = __ldg(xyz + (size_t)j * 3 + 1); o[2] = __ldg(xyz + (size_t)j * 3 + 2);
      }
    }
    return;
  }

  if (CS > 1) {
    if (tid == 0) {
      mbar_init(smem_u32(&s_mbar[0]), 1u);  // one local arming arrive per round; peers complete the tx bytes
      mbar_init(smem_u32(&s_mbar[1]), 1u);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_sync_all();  // barriers initialised + every CTA of the cluster is resident before any DSMEM access
  }

  float px[PPT], py[PPT], pz[PPT], td[PPT];
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    const int k = g + i * T;
    float x = 0.f, y = 0.f, z = 0.f, t = -INFINITY;
    if (k < N) {
      x = __ldg(xyz + (size_t)k * 3 + 0);
      y = __ldg(xyz + (size_t)k * 3 + 1);
      z = __ldg(xyz + (size_t)k * 3 + 2);
      float mag = __fmul_rn(y, y);
      mag = __fmaf_rn(x, x, mag);
      mag = __fmaf_rn(z, z, mag);
      // reference :101 `if (mag <= 1e-3) continue;` -- float promoted to double against a double literal.
      // A skipped point is never a candidate: td = -inf makes fminf(d, td) = -inf, never > best (-1).
      t = ((double)mag <= 1e-3) ? -INFINITY : 1e10f;
    }
    px[i] = x; py[i] = y; pz[i] = z; td[i] = t;
    s_pts[i * FPS_THREADS + tid] = make_float4(x, y, z, 0.f);
  }
  const float p0x = __ldg(xyz + 0), p0y = __ldg(xyz + 1), p0z = __ldg(xyz + 2);
  float x1 = p0x, y1 = p0y, z1 = p0z;  // old = 0
  float *__restrict__ nx_out = new_xyz_all ? new_xyz_all + (size_t)scene * m * 3 : nullptr;
  if (g == 0 && m > 0) {
    idx_out[0] = 0;
    if (nx_out) { nx_out[0] = p0x; nx_out[1] = p0y; nx_out[2] = p0z; }
  }
  uint32_t phases = 0;
  const uint32_t bs_mask = (1u << bs_log2) - 1u;

  for (int j = 1; j < m; ++j) {
    const int par = j & 1;
    // arm this round's barrier: CS records of 32 bytes will land in s_cta[par] (peers may already be sending:
    // the pending arrival keeps the phase open until this arm has happened)
    if (CS > 1 && tid == 0) mbar_arrive_expect_tx_local(smem_u32(&s_mbar[par]), (uint32_t)CS * 32u);
    float best = -1.f;
    int besti = 0;
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      const float d = sqdist_yxz(px[i] - x1, py[i] - y1, pz[i] - z1);
      const float d2 = fminf(d, td[i]);
      td[i] = d2;
      const bool gt = d2 > best;
      besti = gt ? i : besti;
      best = gt ? d2 : best;
    }
    uint32_t hi = 0u, lo = 0u;
    const int k = g + besti * T;
    if (best >= 0.f) {
      const uint32_t slot = (uint32_t)k & bs_mask;
      const uint32_t rev = bs_log2 ? (__brev(slot) >> (32 - bs_log2)) : 0u;
      const uint32_t rk = rev * (uint32_t)Q + ((uint32_t)k >> bs_log2);
      hi = __float_as_uint(best);
      lo = 0xffffffffu - rk;
    }
    {
      const uint32_t hmax = __reduce_max_sync(0xffffffffu, hi);
      const uint32_t lmax = __reduce_max_sync(0xffffffffu, hi == hmax ? lo : 0u);
      const bool none = (hmax | lmax) == 0u;
      const bool win = none ? (lane == 0) : (hi == hmax && lo == lmax);
      if (win) {
        const float4 p = s_pts[besti * FPS_THREADS + tid];
        uint4 *dst = reinterpret_cast<uint4 *>(&s_warp[par][warp]);
        dst[0] = make_uint4(hmax, lmax, (uint32_t)k, 0u);
        reinterpret_cast<float4 *>(dst)[1] = p;
      }
    }
    __syncthreads();
    FpsRec r;
    if (CS == 1) {
      r = fps_pick(s_warp[par], FPS_WARPS, lane);
    } else {
      if (warp == 0) {
        const FpsRec c = fps_pick(s_warp[par], FPS_WARPS, lane);
        if (lane < CS) {
          const uint32_t dst = mapa_u32(smem_u32(&s_cta[par][rank]), (uint32_t)lane);
          const uint32_t bar = mapa_u32(smem_u32(&s_mbar[par]), (uint32_t)lane);
          st_async_v4(dst, c.hi, c.lo, (uint32_t)c.k, 0u, bar);
          st_async_v4(dst + 16, __float_as_uint(c.x), __float_as_uint(c.y), __float_as_uint(c.z), 0u, bar);
        }
      }
      mbar_wait_acquire_cluster(smem_u32(&s_mbar[par]), (phases >> par) & 1u);
      phases ^= (1u << par);
      r = fps_pick(s_cta[par], CS, lane);
    }
    int old;
    if ((r.hi | r.lo) == 0u) {  // no candidate anywhere: reference leaves besti = 0 in every thread
      old = 0; x1 = p0x; y1 = p0y; z1 = p0z;
    } else {
      old = r.k; x1 = r.x; y1 = r.y; z1 = r.z;
    }
    if (g == 0) {
      idx_out[j] = old;
      // the sampled coordinates are already in registers: emitting them here replaces the reference's
      // transpose + gather_points + transpose (pointnet2_modules.py:219-226) with three stores per round
      if (nx_out) { nx_out[j * 3 + 0] = x1; nx_out[j * 3 + 1] = y1; nx_out[j * 3 + 2] = z1; }
    }
  }
  if (CS > 1) cluster_sync_all();  // no CTA may exit while a peer can still write into its shared memory
}

template <int PPT, int FPS_THREADS>
static int launch_fps(const float *xyz, int B, int N, int m, int bs_log2, int Q, int CS, int *idx, float *new_xyz,
                      const int *prefix_flag, cudaStream_t stream, bool probe_only, int *max_clusters) {
  auto kern = fps_kernel<PPT, FPS_THREADS>;
  const size_t smem = (size_t)PPT * FPS_THREADS * sizeof(float4);
  RFD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "fps attr smem");
  if (CS > 8)
    RFD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1), "fps attr cluster");
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(B * CS));
  cfg.blockDim = dim3(FPS_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (probe_only) {
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    if (e != cudaSuccess) { (void)cudaGetLastError(); n = 0; }
    *max_clusters = n;
    return RFD_OK;
  }
  RFD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, xyz, N, m, bs_log2, Q, CS, idx, new_xyz, prefix_flag), "fps launch");
  RFD_CHECK_LAUNCH("fps_kernel");
  return RFD_OK;
}

static int dispatch_fps(int threads, int ppt, const float *xyz, int B, int N, int m, int bs_log2, int Q, int CS,
                        int *idx, float *new_xyz, const int *prefix_flag, cudaStream_t stream, bool probe, int *maxc) {
#define RFD_FPS_CASE(P, T) \
  if (ppt <= P) return launch_fps<P, T>(xyz, B, N, m, bs_log2, Q, CS, idx, new_xyz, prefix_flag, stream, probe, maxc);
  if (threads == 1024) {
    RFD_FPS_CASE(1, 1024) RFD_FPS_CASE(2, 1024) RFD_FPS_CASE(3, 1024) RFD_FPS_CASE(4, 1024) RFD_FPS_CASE(5, 1024)
    RFD_FPS_CASE(6, 1024) RFD_FPS_CASE(8, 1024)
    return RFD_ERR_UNSUPPORTED_SIZE;
  }
  RFD_FPS_CASE(1, 512) RFD_FPS_CASE(2, 512) RFD_FPS_CASE(3, 512) RFD_FPS_CASE(4, 512) RFD_FPS_CASE(6, 512)
  RFD_FPS_CASE(8, 512) RFD_FPS_CASE(10, 512) RFD_FPS_CASE(12, 512) RFD_FPS_CASE(16, 512) RFD_FPS_CASE(20, 512)
  RFD_FPS_CASE(24, 512)
#undef RFD_FPS_CASE
  return RFD_ERR_UNSUPPORTED_SIZE;
}

// ------------------------------------------------------------------------------------------------
// "FPS of an FPS-ordered prefix is the identity" (pointnet2backbone.py:104-113): SA(k+1) samples the points SA(k)
// sampled, in the order SA(k) sampled them, and -- barring ties -- gets back 0, 1, ..., m-1 after m-1 SERIAL rounds.
// Instead of trusting that, it is PROVED per scene, in parallel, with the sampler's own arithmetic:
//   FPS(xyz)[0..m) == (0, 1, ..., m-1)   <=>   for every round j in 1..m-1, point j is the arg-max of
//       D_j(k) = min(1e10, min_{i<j} d(p_k, p_i))   over the non-skipped points k,
//   i.e. for every k != j:  D_j(k) < D_j(j)  or  (D_j(k) == D_j(j) and rank(j) < rank(k))      (tie-break of the sampler).
// own[j] = D_j(j) costs thread j a loop over i < j (pass 1); thread k then walks j = 1..m-1 keeping D_j(k) as a running
// minimum and compares it with own[j] (pass 2): the same n*m distance evaluations as the sampler, but as one wide
// parallel sweep (~10 us) instead of m-1 dependent rounds (0.45 us each).  Any violated comparison (ties resolved the
// other way, NaNs, skipped points among the first m) clears the scene's flag and the full sampler runs for that scene.
constexpr int PFX_THREADS = 128;
constexpr int PFX_MAX_M = 3072;  // pass 2 keeps m float4 records in dynamic shared memory (48 KB without opt-in)

__device__ __forceinline__ bool fps_skipped(float x, float y, float z) {
  float mag = __fmul_rn(y, y);
  mag = __fmaf_rn(x, x, mag);
  mag = __fmaf_rn(z, z, mag);
  return (double)mag <= 1e-3;
}

// pass 1: own[j] = D_j(j) for j in 1..m-1 (-inf if point j is skipped: it can never be sampled); flag[b] = 1
__global__ void __launch_bounds__(PFX_THREADS)
fps_prefix_own_kernel(const float *__restrict__ xyz_all, int N, int m, float *__restrict__ own_all, int *__restrict__ flag) {
  const int b = blockIdx.y;
  const float *__restrict__ xyz = xyz_all + (size_t)b * N * 3;
  const int j = blockIdx.x * PFX_THREADS + threadIdx.x;
  if (j == 0) flag[b] = 1;
  if (j < 1 || j >= m) return;
  const float x = __ldg(xyz + (size_t)j * 3), y = __ldg(xyz + (size_t)j * 3 + 1), z = __ldg(xyz + (size_t)j * 3 + 2);
  float D = fps_skipped(x, y, z) ? -INFINITY : 1e10f;
  for (int i = 0; i < j; ++i) {
    const float d = sqdist_yxz(x - __ldg(xyz + (size_t)i * 3), y - __ldg(xyz + (size_t)i * 3 + 1), z - __ldg(xyz + (size_t)i * 3 + 2));
    D = fminf(d, D);
  }
  own_all[(size_t)b * m + j] = D;
}

// pass 2: thread k checks D_j(k) against own[j] for every round j
__global__ void __launch_bounds__(PFX_THREADS)
fps_prefix_verify_kernel(const float *__restrict__ xyz_all, int N, int m, int bs_log2, int Q,
                         const float *__restrict__ own_all, int *__restrict__ flag) {
  extern __shared__ float4 s_rec[];  // [j] = (p_{j-1}.xyz, own[j]), j = 1..m-1
  const int b = blockIdx.y;
  const float *__restrict__ xyz = xyz_all + (size_t)b * N * 3;
  const float *__restrict__ own = own_all + (size_t)b * m;
  for (int j = 1 + threadIdx.x; j < m; j += PFX_THREADS)
    s_rec[j] = make_float4(__ldg(xyz + (size_t)(j - 1) * 3), __ldg(xyz + (size_t)(j - 1) * 3 + 1),
                           __ldg(xyz + (size_t)(j - 1) * 3 + 2), __ldg(own + j));
  __syncthreads();
  const int k = blockIdx.x * PFX_THREADS + threadIdx.x;
  if (k >= N) return;
  const float x = __ldg(xyz + (size_t)k * 3), y = __ldg(xyz + (size_t)k * 3 + 1), z = __ldg(xyz + (size_t)k * 3 + 2);
  if (fps_skipped(x, y, z)) return;  // never a candidate
  const uint32_t bs_mask = (1u << bs_log2) - 1u;
  auto rank_of = [&](uint32_t p) {
    const uint32_t slot = p & bs_mask;
    const uint32_t rev = bs_log2 ? (__brev(slot) >> (32 - bs_log2)) : 0u;
    return rev * (uint32_t)Q + (p >> bs_log2);
  };
  const uint32_t rk = rank_of((uint32_t)k);
  float D = 1e10f;
  bool ok = true;
#pragma unroll 4
  for (int j = 1; j < m; ++j) {
    const float4 r = s_rec[j];
    const float d = sqdist_yxz(x - r.x, y - r.y, z - r.z);
    D = fminf(d, D);
    if (j != k) {
      const bool fine = (D < r.w) || (D == r.w && rank_of((uint32_t)j) < rk);
      ok = ok && fine;
    }
  }
  if (!ok) flag[b] = 0;
}

// reference cuda_utils.h:15-19 (evaluated with the same double arithmetic)
static int ref_opt_n_threads(int work_size) {
  const int pow_2 = (int)(std::log((double)work_size) / std::log(2.0));
  int v = 1 << pow_2;
  if (v > 512) v = 512;
  if (v < 1) v = 1;
  return v;
}

int fps_plan(int N, int B, int num_sms, int threads, int *cs_out, int *ppt_out) {
  const int max_ppt = threads == 1024 ? FPS_MAX_PPT_1K : FPS_MAX_PPT;
  const int tgt = threads == 1024 ? 5 : 10;                 // points per thread aimed for on big clouds
  const int need = (N + threads - 1) / threads;             // point slots per thread column
  int cs = 1;
  if (N > 8192) {
    while (cs < FPS_MAX_CS && (need + cs - 1) / cs > tgt) cs *= 2;
    // oversubscribed GPU: prefer fewer, fatter CTAs per scene as long as the points still fit in registers
    while (cs > 1 && (long long)B * cs > num_sms && (need + cs / 2 - 1) / (cs / 2) <= max_ppt) cs /= 2;
  }
  while (cs < FPS_MAX_CS && (need + cs - 1) / cs > max_ppt) cs *= 2;
  const int ppt = (need + cs - 1) / cs;
  if (ppt > max_ppt) return RFD_ERR_UNSUPPORTED_SIZE;
  *cs_out = cs;
  *ppt_out = ppt;
  return RFD_OK;
}

}  // namespace rfd

extern "C" int rfd_furthest_point_sampling(const float *xyz, int B, int N, int m, int *idx, void *stream) {
  return rfd_furthest_point_sampling_cond(xyz, B, N, m, nullptr, idx, nullptr, stream);
}

extern "C" int rfd_furthest_point_sampling_xyz(const float *xyz, int B, int N, int m, int *idx, float *new_xyz,
                                               void *stream) {
  return rfd_furthest_point_sampling_cond(xyz, B, N, m, nullptr, idx, new_xyz, stream);
}

extern "C" int rfd_furthest_point_sampling_cond(const float *xyz, int B, int N, int m, const int *prefix_flag, int *idx,
                                                float *new_xyz, void *stream) {
  using namespace rfd;
  if (B < 0 || N < 1 || m < 0 || (B > 0 && (!xyz || (m > 0 && !idx)))) return RFD_ERR_INVALID_ARGUMENT;
  if (B == 0 || m == 0) return RFD_OK;
  int dev = 0, sms = 148;
  RFD_CHECK_CUDA(cudaGetDevice(&dev), "fps getdevice");
  RFD_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "fps sms");
  // CTA width: 512 threads.  1024-thread CTAs (5 points/thread, 8 warps per scheduler) were measured SLOWER on the
  // 80k-point layer (2.28 ms vs 1.84 ms for 4 scenes): the round is bound by the reduction / exchange chain, whose
  // cost grows with the warp count, not by the register update.  RFD_FPS_THREADS=1024 selects them for experiments.
  int threads = 512;
  if (const char *e = getenv("RFD_FPS_THREADS")) { const int v = atoi(e); if (v == 512 || v == 1024) threads = v; }
  int cs = 1, ppt = 1;
  int rc = fps_plan(N, B, sms, threads, &cs, &ppt);
  if (rc != RFD_OK && threads == 1024) { threads = 512; rc = fps_plan(N, B, sms, threads, &cs, &ppt); }
  if (rc != RFD_OK) return rc;
  const int bs = ref_opt_n_threads(N);
  int bs_log2 = 0;
  while ((1 << bs_log2) < bs) ++bs_log2;
  const int Q = (N + bs - 1) / bs;
  cudaStream_t st = as_stream(stream);
  // Choose the cluster size by a small cost model: a round costs ~0.55 us (reduction + DSMEM exchange chain) plus
  // ~0.034 us per point held by a thread (measured, profiles/); clusters that cannot be co-resident run in waves.
  // Few GPCs can host a 16-CTA cluster at once (cudaOccupancyMaxActiveClusters), so with many scenes 8-CTA clusters
  // holding 20 points per thread finish earlier than two waves of 16-CTA clusters.
  if (cs > 1) {
    const int max_ppt = threads == 1024 ? FPS_MAX_PPT_1K : FPS_MAX_PPT;
    const int need = (N + threads - 1) / threads;
    double best_cost = 1e30;
    int best_cs = 0, best_ppt = 0;
    for (int c = cs; c >= 2 && c >= cs / 4; c /= 2) {
      const int p = (need + c - 1) / c;
      if (p > max_ppt) break;
      // [threads==1024][ppt][cs] -> max active clusters + 1 (0 = unknown); relaxed atomics: concurrent callers may both
      // probe and store the same value
      static std::atomic<int> cache[2][32][17];
      std::atomic<int> &slot = cache[threads == 1024][p][c];
      int known = slot.load(std::memory_order_relaxed);
      if (known == 0) {
        int probed = 0;
        rc = dispatch_fps(threads, p, xyz, B, N, m, bs_log2, Q, c, idx, new_xyz, nullptr, st, true, &probed);
        if (rc != RFD_OK) return rc;
        known = probed + 1;
        slot.store(known, std::memory_order_relaxed);
      }
      const int maxc = known - 1;
      if (getenv("RFD_FPS_DEBUG")) fprintf(stderr, "[rfd fps] candidate cluster=%d ppt=%d max active clusters=%d\n", c, p, maxc);
      if (maxc <= 0) continue;  // e.g. a 16-CTA (non-portable) cluster is not schedulable on this part
      const int waves = (B + maxc - 1) / maxc;
      const double cost = waves * (0.55 + 0.034 * p);
      if (cost < best_cost) { best_cost = cost; best_cs = c; best_ppt = p; }
    }
    if (best_cs == 0) {  // nothing schedulable as a cluster: one fat CTA per scene if the points fit
      best_cs = 1;
      best_ppt = need;
      if (best_ppt > max_ppt) return RFD_ERR_UNSUPPORTED_SIZE;
    }
    cs = best_cs;
    ppt = best_ppt;
  }
  if (getenv("RFD_FPS_DEBUG"))
    fprintf(stderr, "[rfd fps] B=%d N=%d m=%d threads=%d cluster=%d points/thread=%d\n", B, N, m, threads, cs, ppt);
  return dispatch_fps(threads, ppt, xyz, B, N, m, bs_log2, Q, cs, idx, new_xyz, prefix_flag, st, false, nullptr);
}

extern "C" int rfd_fps_prefix_check(const float *xyz, int B, int N, int m, float *own_ws, int *flag, void *stream) {
  using namespace rfd;
  if (B < 0 || N < 1 || m < 0 || (B > 0 && (!xyz || !flag || (m > 1 && !own_ws)))) return RFD_ERR_INVALID_ARGUMENT;
  if (B == 0) return RFD_OK;
  cudaStream_t st = as_stream(stream);
  if (m > N || m > PFX_MAX_M || B > 65535) {  // m > N can never be the identity; larger m: not supported, run the full FPS
    RFD_CHECK_CUDA(cudaMemsetAsync(flag, 0, sizeof(int) * (size_t)B, st), "fps_prefix memset");
    return RFD_OK;
  }
  const int bs = ref_opt_n_threads(N);
  int bs_log2 = 0;
  while ((1 << bs_log2) < bs) ++bs_log2;
  const int Q = (N + bs - 1) / bs;
  fps_prefix_own_kernel<<<dim3(h_ceil_div(m > 0 ? m : 1, PFX_THREADS), B), PFX_THREADS, 0, st>>>(xyz, N, m, own_ws, flag);
  RFD_CHECK_LAUNCH("fps_prefix_own_kernel");
  if (m > 1) {
    fps_prefix_verify_kernel<<<dim3(h_ceil_div(N, PFX_THREADS), B), PFX_THREADS, (size_t)m * sizeof(float4), st>>>(
        xyz, N, m, bs_log2, Q, own_ws, flag);
    RFD_CHECK_LAUNCH("fps_prefix_verify_kernel");
  }
  return RFD_OK;
}
