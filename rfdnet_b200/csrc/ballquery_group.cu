// ballquery_group.cu -- ball query, grouping gather, their fusion, point gather, and the scatter-add grads.
//
// Reference kernels replaced (paths relative to _ext-src/src/):
//   query_ball_point_kernel   ball_query_gpu.cu:9-44     one CTA per scene, one THREAD per query scanning all N
//   group_points_kernel       group_points_gpu.cu:8-28   one CTA per scene, 4-byte random gathers
//   group_points_grad_kernel  group_points_gpu.cu:43-64
//   gather_points_kernel      sampling_gpu.cu:8-20 ; gather_points_grad_kernel :34-47
//   QueryAndGroup.forward     pointnet2_utils.py:302-361  (ball_query + group(xyz) + sub + div + group(feat) + cat)
//
// Design of the fused kernel (rfd_query_and_group), an HBM-write-bound operator (the (B,3+C,M,S) output is 10-20x
// larger than everything it reads):
//   phase 1  ball query, one WARP per query: the 32 lanes test 32 candidates per step, a ballot + popc prefix assigns
//            output slots in ascending index order -- exactly the "first nsample hits by ascending k" semantics of the
//            reference -- and the warp stops as soon as nsample hits are found.  Candidates come from
//              * the scene's cloud staged in shared memory by one bulk (TMA) copy per 4096-point chunk (N < 8192), or
//              * a per-scene uniform grid (N >= 8192): points sorted by cell as float4 (x,y,z,id) by ONE cluster
//                kernel; a query reads the 9 contiguous z-runs of its 3x3x3 neighbourhood with coalesced 16-byte loads.
//            The indices of a CTA's queries (<= 2048 (query,sample) slots) stay in shared memory.
//   phase 2  relative xyz channels: one thread per slot, (p - c) * (1/r) computed on the fly, 128-byte coalesced stores.
//   phase 3  feature channels.  Gathering 4-byte elements from the channel-major (B,C,N) tensor costs one L1 wavefront
//            per element; instead the features are first transposed to point-major (B,N,C) (a copy of the SMALL
//            tensor, 1/16 of the output), so a (query,sample) slot is one contiguous row: a warp reads 32 slots x 32
//            channels with eight 16-byte loads per lane (4 rows x 128 B per instruction), transposes the 4-KB tile
//            through XOR-swizzled shared memory (conflict-free 16-byte stores and loads) and writes each channel's
//            32 consecutive slots as one 128-byte coalesced store.  No per-element index arithmetic, no div/mod.
// None of the reference's four intermediate passes over the grouped tensor exists.
#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"
#include "umma.cuh"

namespace rfd {

constexpr int BQ_WARPS = 8;  // warps per CTA (stand-alone ball query: queries per CTA)
constexpr int BQ_THREADS = BQ_WARPS * 32;
constexpr int QG_MAX_S = 1024;      // nsample limit of the fused kernel
constexpr int QG_MAX_SLOTS = 2048;  // (query, sample) slots per CTA
constexpr int QG_CHUNK = 4096;      // points staged in shared memory per pass (48 KB)
constexpr int QG_TILE = 4096;       // per-warp transposition tile: 32 slots x 32 channels fp32

// Scan of a global-memory cloud for one query by one warp.  Hits are reported through `emit(pos, k)` in ascending k,
// pos < nsample.  Returns the hit count (capped at nsample) and the first hit index.
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

// Same scan over `np` points staged in shared memory as xyz triples (conflict-free stride-3 reads); `k0` = index of the
// first staged point; continues a scan that already has `cnt` hits.
template <typename Emit>
__device__ __forceinline__ int ball_scan_smem(const float *s_xyz, int np, int k0, float qx, float qy, float qz,
                                              float radius2, int nsample, int lane, int cnt, int &first, Emit emit) {
  for (int base = 0; base < np && cnt < nsample; base += 32) {
    const int i = base + lane;
    bool hit = false;
    if (i < np) {
      const float x = s_xyz[3 * i], y = s_xyz[3 * i + 1], z = s_xyz[3 * i + 2];
      hit = sqdist_yxz(qx - x, qy - y, qz - z) < radius2;
    }
    const unsigned mask = __ballot_sync(0xffffffffu, hit);
    if (mask) {
      if (cnt == 0) first = k0 + base + __ffs(mask) - 1;
      const int pos = cnt + __popc(mask & ((1u << lane) - 1u));
      if (hit && pos < nsample) emit(pos, k0 + i);
      cnt += __popc(mask);
    }
  }
  return cnt;
}

// ------------------------------------------------------------------------------------------------
// Uniform-grid candidate search for large clouds (SA1: N = 80000, 2048 queries -> 164 M brute-force pair tests).
// Points are binned into cells of edge >= 1.001 * radius, so every point within `radius` of a query lies in the
// 3x3x3 cell neighbourhood of the query's cell.  The hit PREDICATE is unchanged (same fp32 arithmetic as the
// reference, on bit-identical copies of the coordinates), only the candidate set shrinks; the reference's "first
// nsample hits by ascending index" order is restored by ranking the collected hits by index (rank = number of hits
// with a smaller index), so the output is bit-identical to the brute-force scan.  A query with more than GRID_CAP hits
// falls back to the brute-force scan.
constexpr int GRID_G = 32;                        // max cells per axis
constexpr int GRID_NC = GRID_G * GRID_G * GRID_G;  // 32768
constexpr int GRID_CAP = 512;                      // hits kept per query before falling back
constexpr int GRID_MIN_N = 8192;                   // use the grid from this cloud size on
constexpr int GB_CS = 8;                           // grid build: CTAs per scene (one thread-block cluster)
constexpr int GB_THREADS = 1024;

struct GridScene {
  float ox, oy, oz, inv_cell;
  int gx, gy, gz, pad;
};

__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// integer cell coordinate along one axis, clamped to [-2, g+1] before the float->int conversion
__device__ __forceinline__ int grid_coord(float v, float o, float inv, int g) {
  float t = floorf((v - o) * inv);
  t = fminf(fmaxf(t, -2.f), (float)(g + 1));
  return (t == t) ? (int)t : 0;
}
__device__ __forceinline__ int grid_cell(const GridScene &g, float x, float y, float z) {
  const int cx = min(max(grid_coord(x, g.ox, g.inv_cell, g.gx), 0), g.gx - 1);
  const int cy = min(max(grid_coord(y, g.oy, g.inv_cell, g.gy), 0), g.gy - 1);
  const int cz = min(max(grid_coord(z, g.oz, g.inv_cell, g.gz), 0), g.gz - 1);
  return (cx * g.gy + cy) * g.gz + cz;
}

__device__ __forceinline__ int ld_shared_cluster_s32(uint32_t cluster_addr) {
  int v;
  asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(cluster_addr));
  return v;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}

// The whole grid build of one scene in ONE kernel, by one cluster of GB_CS CTAs (replaces the six launches of round 1):
//   0  bounding box: CTA-local reduction, exchanged through distributed shared memory -> every CTA derives the same grid
//   1  cell histogram: atomicAdd on count[] (global; all-zero on entry)
//   2  exclusive scan of the histogram into shared memory (every CTA; CTA 0 also publishes it as start[])
//   3  scatter: slot = start[cell] + (atomicSub(count[cell]) - 1); sorted[slot] = (x, y, z, id)
// count[] is all-zero again on exit (every increment is undone by one decrement), so the workspace needs no memset
// between calls.  Order inside a cell is arbitrary; the query ranks its hits by id.
__global__ void __cluster_dims__(GB_CS, 1, 1) __launch_bounds__(GB_THREADS, 1)
grid_build_kernel(const float *__restrict__ xyz, int n, float radius, GridScene *__restrict__ gs,
                  int *__restrict__ count, int *__restrict__ start, float4 *__restrict__ sorted) {
  extern __shared__ int s_start[];  // GRID_NC
  __shared__ int s_bbox[6], s_box[6];
  __shared__ int s_wsum[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  const uint32_t rank = umma::cluster_ctarank();
  const float *p = xyz + (size_t)b * n * 3;
  count += (size_t)b * (GRID_NC + 1);
  start += (size_t)b * (GRID_NC + 1);
  sorted += (size_t)b * n;
  const int stride = GB_CS * GB_THREADS;
  // ---- 0: bounding box
  if (tid < 6) s_bbox[tid] = tid < 3 ? 0x7fffffff : (int)0x80000000;
  __syncthreads();
  {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int k = rank * GB_THREADS + tid; k < n; k += stride) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float v = __ldg(p + (size_t)k * 3 + c);
        if (v == v) { mn[c] = fminf(mn[c], v); mx[c] = fmaxf(mx[c], v); }
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], off));
        mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], off));
      }
      if (lane == 0) {
        atomicMin(&s_bbox[c], f2ord(mn[c]));
        atomicMax(&s_bbox[3 + c], f2ord(mx[c]));
      }
    }
  }
  umma::cluster_sync();
  if (tid < 6) {
    int v = s_bbox[tid];
    const uint32_t a = umma::smem_u32(&s_bbox[tid]);
    for (uint32_t r = 0; r < GB_CS; ++r) {
      const int o = ld_shared_cluster_s32(mapa_shared(a, r));
      v = tid < 3 ? min(v, o) : max(v, o);
    }
    s_box[tid] = v;
  }
  __syncthreads();
  GridScene g;
  {
    float mn[3], ext[3], emax = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      mn[c] = ord2f(s_box[c]);
      const float mx = ord2f(s_box[3 + c]);
      ext[c] = (mx >= mn[c]) ? mx - mn[c] : 0.f;
      if (!(ext[c] < 3.0e38f)) ext[c] = 3.0e38f;
      emax = fmaxf(emax, ext[c]);
    }
    float cell = fmaxf(radius * 1.001f, emax / (float)(GRID_G - 1) * 1.0001f);
    if (!(cell > 0.f)) cell = 1.f;
    g.ox = mn[0]; g.oy = mn[1]; g.oz = mn[2];
    g.inv_cell = 1.0f / cell;
    g.gx = min(GRID_G, (int)(ext[0] * g.inv_cell) + 1);
    g.gy = min(GRID_G, (int)(ext[1] * g.inv_cell) + 1);
    g.gz = min(GRID_G, (int)(ext[2] * g.inv_cell) + 1);
    g.pad = 0;
  }
  if (rank == 0 && tid == 0) gs[b] = g;
  // ---- 1: histogram
  for (int k = rank * GB_THREADS + tid; k < n; k += stride)
    atomicAdd(count + grid_cell(g, __ldg(p + (size_t)k * 3), __ldg(p + (size_t)k * 3 + 1), __ldg(p + (size_t)k * 3 + 2)), 1);
  umma::cluster_sync();
  // ---- 2: exclusive scan (warp w owns cells [1024 w, 1024 w + 1024))
  const int ncell = g.gx * g.gy * g.gz;
  {
    int run = 0;
    const int c0 = warp * (GRID_NC / 32);
    if (c0 < ncell) {
#pragma unroll 4
      for (int it = 0; it < GRID_NC / 32 / 32; ++it) {
        const int c = c0 + it * 32 + lane;
        const int v = c < ncell ? __ldcg(count + c) : 0;
        int inc = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, inc, off);
          if (lane >= off) inc += t;
        }
        s_start[c] = run + inc - v;
        run += __shfl_sync(0xffffffffu, inc, 31);
      }
    }
    if (lane == 0) s_wsum[warp] = run;
    __syncthreads();
    if (warp == 0) {
      const int v = s_wsum[lane];
      int inc = v;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += t;
      }
      s_wsum[lane] = inc - v;
    }
    __syncthreads();
    const int woff = s_wsum[warp];
    if (c0 < ncell) {
      for (int it = 0; it < GRID_NC / 32 / 32; ++it) {
        const int c = c0 + it * 32 + lane;
        const int v = s_start[c] + woff;
        s_start[c] = v;
        if (rank == 0 && c <= ncell) start[c] = c < ncell ? v : n;
      }
    }
    if (rank == 0 && tid == 0 && (ncell & 1023) == 0) start[ncell] = n;  // ncell on a warp-segment boundary
  }
  umma::cluster_sync();  // every CTA has read the histogram before anyone starts undoing it
  // ---- 3: scatter
  for (int k = rank * GB_THREADS + tid; k < n; k += stride) {
    const float x = __ldg(p + (size_t)k * 3), y = __ldg(p + (size_t)k * 3 + 1), z = __ldg(p + (size_t)k * 3 + 2);
    const int c = grid_cell(g, x, y, z);
    const int pos = s_start[c] + atomicSub(count + c, 1) - 1;
    sorted[pos] = make_float4(x, y, z, __int_as_float(k));
  }
}

// Grid scan for one query by one warp: collects the hits of the 27 neighbouring cells in `hits` (shared, GRID_CAP
// ints), ranks them by index and writes the first nsample into row[] in ascending index order.
// Returns the hit count capped at nsample, or -1 if more than GRID_CAP hits were found (caller falls back).
__device__ __forceinline__ int ball_scan_grid(const float4 *__restrict__ sorted, const GridScene &g,
                                              const int *__restrict__ start, float qx, float qy, float qz,
                                              float radius2, int nsample, int lane, int *hits, int *row, int &first) {
  const int cx = grid_coord(qx, g.ox, g.inv_cell, g.gx), cy = grid_coord(qy, g.oy, g.inv_cell, g.gy),
            cz = grid_coord(qz, g.oz, g.inv_cell, g.gz);
  int H = 0;
  for (int ix = max(cx - 1, 0); ix <= min(cx + 1, g.gx - 1); ++ix)
    for (int iy = max(cy - 1, 0); iy <= min(cy + 1, g.gy - 1); ++iy) {
      // cells (ix, iy, z0..z1) are contiguous in the cell order => one contiguous run of the sorted array
      const int z0 = max(cz - 1, 0), z1 = min(cz + 1, g.gz - 1);
      if (z0 > z1) continue;
      const int c0 = (ix * g.gy + iy) * g.gz;
      const int beg = __ldg(start + c0 + z0), end = __ldg(start + c0 + z1 + 1);
      for (int base = beg; base < end; base += 32) {
        const int e = base + lane;
        bool hit = false;
        int k = 0;
        if (e < end) {
          const float4 pt = __ldg(sorted + e);
          k = __float_as_int(pt.w);
          hit = sqdist_yxz(qx - pt.x, qy - pt.y, qz - pt.z) < radius2;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, hit);
        if (mask) {
          const int pos = H + __popc(mask & ((1u << lane) - 1u));
          if (hit && pos < GRID_CAP) hits[pos] = k;
          H += __popc(mask);
        }
      }
    }
  if (H > GRID_CAP) return -1;
  __syncwarp();
  int mn = 0x7fffffff;
  for (int h = lane; h < H; h += 32) {
    const int id = hits[h];
    int rank = 0;
    for (int t = 0; t < H; ++t) rank += (hits[t] < id);
    if (rank < nsample) row[rank] = id;
    mn = min(mn, id);
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, off));
  first = H ? mn : 0;
  __syncwarp();
  return H < nsample ? H : nsample;
}

__global__ void __launch_bounds__(BQ_THREADS)
ball_query_kernel(const float *__restrict__ new_xyz, const float *__restrict__ xyz, int n, int m, float radius,
                  int nsample, int *__restrict__ idx, const GridScene *__restrict__ gs, const int *__restrict__ gstart,
                  const float4 *__restrict__ gsorted) {
  __shared__ int s_hits[BQ_WARPS][GRID_CAP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int j = blockIdx.x * BQ_WARPS + warp;
  if (j >= m) return;
  xyz += (size_t)b * n * 3;
  new_xyz += (size_t)b * m * 3;
  int *row = idx + ((size_t)b * m + j) * nsample;
  const float radius2 = __fmul_rn(radius, radius);  // reference :22
  const float qx = __ldg(new_xyz + j * 3 + 0), qy = __ldg(new_xyz + j * 3 + 1), qz = __ldg(new_xyz + j * 3 + 2);
  int first, cnt = -1;
  if (gs)
    cnt = ball_scan_grid(gsorted + (size_t)b * n, gs[b], gstart + (size_t)b * (GRID_NC + 1), qx, qy, qz, radius2,
                         nsample, lane, s_hits[warp], row, first);
  if (cnt < 0)
    cnt = ball_scan(xyz, n, qx, qy, qz, radius2, nsample, lane, first, [&](int pos, int k) { row[pos] = k; });
  // reference :35-39: the first hit pre-fills every slot; no hit at all leaves the zero-initialised row
  const int fill = cnt == 0 ? 0 : first;
  for (int l = cnt + lane; l < nsample; l += 32) row[l] = fill;
}

// features (B,C,N) -> point-major (B,N,Cp), Cp = C rounded up to 4 (zero padded).  grid (ceil(N/32), ceil(Cp/32), B)
__global__ void __launch_bounds__(256)
transpose_features_kernel(const float *__restrict__ f, int C, int N, int Cp, float *__restrict__ ft) {
  __shared__ float t[32][33];
  const int b = blockIdx.z, n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  f += (size_t)b * C * N;
  ft += (size_t)b * N * Cp;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i, n = n0 + tx;
    t[ty + 8 * i][tx] = (c < C && n < N) ? __ldg(f + (size_t)c * N + n) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty + 8 * i, c = c0 + tx;
    if (n < N && c < Cp) ft[(size_t)n * Cp + c] = t[tx][ty + 8 * i];
  }
}

struct QgParams {
  const float *xyz, *new_xyz, *features, *feat_t;  // feat_t: point-major copy (B,N,Cp) or nullptr (direct gathers)
  int n, m, C, Cp, nsample, qt;                    // qt: queries per CTA
  float radius;
  int use_xyz, normalize_xyz, s_shift;             // s_shift: log2(nsample) or -1
  float *new_features, *grouped_xyz;
  int *idx_out;
  const GridScene *gs;
  const int *gstart;
  const float4 *gsorted;
  int cloud_floats;                                // shared-memory staging area (floats), 0 in grid mode
};

// fused ball query + group.  grid = (ceil(M/qt), B), 256 threads, dynamic shared memory:
//   [s_idx: qt*S ints][s_q: qt*3 floats][s_cnt, s_first: qt ints each][mbarrier 16 B]
//   [cloud: cloud_floats floats | hits: 8 x GRID_CAP ints][tiles: 8 x 4 KB (only with feat_t)]
__global__ void __launch_bounds__(BQ_THREADS)
query_and_group_kernel(const QgParams P) {
  extern __shared__ __align__(128) uint8_t qg_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int S = P.nsample, n = P.n, m = P.m;
  const int j0 = blockIdx.x * P.qt;
  const int nq = min(P.qt, m - j0);  // queries of this CTA
  const int nslots = nq * S;
  int *s_idx = reinterpret_cast<int *>(qg_smem);
  float *s_q = reinterpret_cast<float *>(s_idx + P.qt * S);
  int *s_cnt = reinterpret_cast<int *>(s_q + P.qt * 3);
  int *s_first = s_cnt + P.qt;
  uint8_t *p8 = reinterpret_cast<uint8_t *>(s_first + P.qt);
  p8 = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(p8) + 15) & ~uintptr_t(15));
  uint64_t *s_bar = reinterpret_cast<uint64_t *>(p8);
  p8 += 16;
  float *s_cloud = reinterpret_cast<float *>(p8);
  int *s_hits = reinterpret_cast<int *>(p8);
  p8 += P.gs ? BQ_WARPS * GRID_CAP * 4 : P.cloud_floats * 4;
  p8 = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(p8) + 127) & ~uintptr_t(127));
  uint8_t *s_tiles = p8;

  const float *xyz = P.xyz + (size_t)b * n * 3;
  const float *new_xyz = P.new_xyz + (size_t)b * m * 3;
  const float radius2 = __fmul_rn(P.radius, P.radius);
  for (int e = threadIdx.x; e < nq * 3; e += BQ_THREADS) s_q[e] = __ldg(new_xyz + (size_t)j0 * 3 + e);
  if (threadIdx.x == 0) { umma::mbar_init(s_bar, 1); umma::fence_barrier_init(); }
  __syncthreads();

  // ---------------- phase 1: ball query -> s_idx
  if (P.gs) {
    const GridScene g = P.gs[b];
    for (int q = warp; q < nq; q += BQ_WARPS) {
      const float qx = s_q[q * 3], qy = s_q[q * 3 + 1], qz = s_q[q * 3 + 2];
      int *row = s_idx + q * S;
      int first = 0, cnt = -1;
      if (S <= GRID_CAP / 2)
        cnt = ball_scan_grid(P.gsorted + (size_t)b * n, g, P.gstart + (size_t)b * (GRID_NC + 1), qx, qy, qz, radius2, S,
                             lane, s_hits + warp * GRID_CAP, row, first);
      if (cnt < 0)
        cnt = ball_scan(xyz, n, qx, qy, qz, radius2, S, lane, first, [&](int pos, int k) { row[pos] = k; });
      const int fill = cnt == 0 ? 0 : first;
      for (int l = cnt + lane; l < S; l += 32) row[l] = fill;
    }
  } else {
    for (int q = threadIdx.x; q < nq; q += BQ_THREADS) { s_cnt[q] = 0; s_first[q] = 0; }
    const int chunk = P.cloud_floats / 3;
    uint32_t parity = 0;
    for (int k0 = 0; k0 < n; k0 += chunk) {
      const int np = min(chunk, n - k0);
      const float *src = xyz + (size_t)k0 * 3;
      // stage the chunk: one bulk (TMA) copy when the source is 16-byte aligned, plus a <= 12-byte tail
      const uint32_t bytes = (uint32_t)np * 12u;
      const bool bulk = (reinterpret_cast<uintptr_t>(src) & 15) == 0 && bytes >= 16;
      const uint32_t bulk_bytes = bulk ? (bytes & ~15u) : 0u;
      __syncthreads();  // previous chunk fully consumed (and s_cnt initialised)
      if (bulk) {
        if (threadIdx.x == 0) {
          umma::fence_proxy_async_smem();  // earlier generic-proxy accesses of the staging area vs the async-proxy write
          umma::mbar_arrive_expect_tx(s_bar, bulk_bytes);
          umma::bulk_g2s(s_cloud, src, bulk_bytes, s_bar);
        }
      }
      for (int e = (int)(bulk_bytes >> 2) + threadIdx.x; e < np * 3; e += BQ_THREADS) s_cloud[e] = __ldg(src + e);
      if (bulk) { umma::mbar_wait(s_bar, parity); parity ^= 1u; }
      __syncthreads();
      for (int q = warp; q < nq; q += BQ_WARPS) {
        int cnt = s_cnt[q], first = s_first[q];
        if (cnt >= S) continue;
        int *row = s_idx + q * S;
        cnt = ball_scan_smem(s_cloud, np, k0, s_q[q * 3], s_q[q * 3 + 1], s_q[q * 3 + 2], radius2, S, lane, cnt, first,
                             [&](int pos, int k) { row[pos] = k; });
        __syncwarp();
        if (lane == 0) { s_cnt[q] = cnt; s_first[q] = first; }
      }
    }
    __syncthreads();
    for (int q = warp; q < nq; q += BQ_WARPS) {
      const int cnt = min(s_cnt[q], S);
      const int fill = cnt == 0 ? 0 : s_first[q];
      for (int l = cnt + lane; l < S; l += 32) s_idx[q * S + l] = fill;
    }
  }
  __syncthreads();

  const size_t MS = (size_t)m * S;
  const size_t slot_base = (size_t)j0 * S;  // first (query, sample) slot of this CTA inside a channel plane
  if (P.idx_out) {
    int *dst = P.idx_out + (size_t)b * MS + slot_base;
    for (int e = threadIdx.x; e < nslots; e += BQ_THREADS) dst[e] = s_idx[e];
  }
  const int cx = P.use_xyz ? 3 : 0;
  const int Ct = cx + P.C;
  float *out = P.new_features + (size_t)b * Ct * MS + slot_base;
  // ---------------- phase 2: relative xyz channels
  if (P.use_xyz || P.grouped_xyz) {
    // torch lowers `tensor /= python_float` on CUDA to a multiply with the f32 reciprocal (pointnet2_utils.py:337)
    const float inv_r = P.normalize_xyz ? __frcp_rn(P.radius) : 1.0f;
    float *gx = P.grouped_xyz ? P.grouped_xyz + (size_t)b * 3 * MS + slot_base : nullptr;
    for (int e = threadIdx.x; e < nslots; e += BQ_THREADS) {
      const int q = P.s_shift >= 0 ? (e >> P.s_shift) : e / S;
      const int k = s_idx[e];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float v = __fsub_rn(__ldg(xyz + (size_t)k * 3 + c), s_q[q * 3 + c]);  // :335 grouped_xyz -= new_xyz
        if (P.normalize_xyz) v = __fmul_rn(v, inv_r);                          // :337
        if (P.use_xyz) out[(size_t)c * MS + e] = v;
        if (gx) gx[(size_t)c * MS + e] = v;
      }
    }
  }
  // ---------------- phase 3: feature channels  out[b][cx+c][j][s] = features[b][c][idx]
  if (P.C > 0) {
    float *o = out + (size_t)cx * MS;
    if (!P.feat_t) {
      const float *f = P.features + (size_t)b * P.C * n;
      for (int e = threadIdx.x; e < nslots; e += BQ_THREADS) {
        const int k = s_idx[e];
        for (int c = 0; c < P.C; ++c) o[(size_t)c * MS + e] = __ldg(f + (size_t)c * n + k);
      }
    } else {
      const int Cp = P.Cp, C = P.C;
      const float *ft = P.feat_t + (size_t)b * n * Cp;
      const uint32_t tile = umma::smem_u32(s_tiles + warp * QG_TILE);
      const int rsub = lane >> 3, jl = lane & 7;
      const int ntile = (nslots + 31) >> 5;
      for (int t = warp; t < ntile; t += BQ_WARPS) {
        const int slot0 = t * 32;
        const float *rowp[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int slot = slot0 + 4 * i + rsub;
          rowp[i] = slot < nslots ? ft + (size_t)s_idx[slot] * Cp + 4 * jl : nullptr;
        }
        const int myslot = slot0 + lane;
        const bool sv = myslot < nslots;
        float *om = o + myslot;
        for (int c0 = 0; c0 < Cp; c0 += 32) {
          const bool cv = c0 + 4 * jl < Cp;
          float4 v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            v[i] = (cv && rowp[i]) ? __ldg(reinterpret_cast<const float4 *>(rowp[i] + c0)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = 4 * i + rsub;
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(tile + r * 128 + ((jl ^ (r & 7)) << 4)),
                         "f"(v[i].x), "f"(v[i].y), "f"(v[i].z), "f"(v[i].w)
                         : "memory");
          }
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 w = umma::lds_f4(tile + lane * 128 + ((j ^ (lane & 7)) << 4));
            const int ch = c0 + 4 * j;
            if (sv) {
              if (ch < C) om[(size_t)ch * MS] = w.x;
              if (ch + 1 < C) om[(size_t)(ch + 1) * MS] = w.y;
              if (ch + 2 < C) om[(size_t)(ch + 2) * MS] = w.z;
              if (ch + 3 < C) om[(size_t)(ch + 3) * MS] = w.w;
            }
          }
          __syncwarp();
        }
      }
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

// ------------------------------------------------------------------------------------------------
// Persistent device workspace, one per (device, stream): the grid (count/start/sorted) and the point-major feature
// copy.  Grown on demand with cudaMalloc (never inside the steady state, so the calls are capturable in a CUDA graph
// after one eager warm-up); `count` is kept all-zero between calls by the build kernel itself.
struct QgWorkspace {
  int *count = nullptr;       // count_scenes x (GRID_NC + 1), all-zero invariant
  int count_scenes = 0;
  uint8_t *scratch = nullptr; // gs | start | sorted | feat_t
  size_t scratch_bytes = 0;
};
static std::mutex g_ws_mu;
static std::map<std::pair<int, cudaStream_t>, QgWorkspace> g_ws;

static int ws_get(cudaStream_t st, int count_scenes, size_t scratch_bytes, QgWorkspace *out) {
  int dev = 0;
  RFD_CHECK_CUDA(cudaGetDevice(&dev), "workspace getdevice");
  std::lock_guard<std::mutex> lk(g_ws_mu);
  QgWorkspace &w = g_ws[std::make_pair(dev, st)];
  if (count_scenes > w.count_scenes) {
    if (w.count) RFD_CHECK_CUDA(cudaFree(w.count), "workspace free");
    w.count = nullptr; w.count_scenes = 0;
    const size_t bytes = sizeof(int) * (size_t)count_scenes * (GRID_NC + 1);
    RFD_CHECK_CUDA(cudaMalloc(&w.count, bytes), "workspace alloc (run once eagerly before CUDA-graph capture)");
    RFD_CHECK_CUDA(cudaMemset(w.count, 0, bytes), "workspace memset");
    w.count_scenes = count_scenes;
  }
  if (scratch_bytes > w.scratch_bytes) {
    if (w.scratch) RFD_CHECK_CUDA(cudaFree(w.scratch), "workspace free");
    w.scratch = nullptr; w.scratch_bytes = 0;
    const size_t bytes = scratch_bytes + (scratch_bytes >> 2);  // headroom: avoid regrowing for slightly larger calls
    RFD_CHECK_CUDA(cudaMalloc(&w.scratch, bytes), "workspace alloc (run once eagerly before CUDA-graph capture)");
    w.scratch_bytes = bytes;
  }
  *out = w;
  return RFD_OK;
}

static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

struct GridPtrs {
  GridScene *gs = nullptr;
  int *start = nullptr;
  float4 *sorted = nullptr;
  float *feat_t = nullptr;
};

// carve the scratch area and (when use_grid) build the per-scene grids: ONE kernel launch
static int prepare(const float *xyz, int B, int N, float radius, bool use_grid, size_t feat_t_bytes, cudaStream_t st,
                   GridPtrs *g) {
  const size_t n_gs = use_grid ? al256(sizeof(GridScene) * (size_t)B) : 0;
  const size_t n_start = use_grid ? al256(sizeof(int) * (size_t)B * (GRID_NC + 1)) : 0;
  const size_t n_sorted = use_grid ? al256(sizeof(float4) * (size_t)B * N) : 0;
  const size_t total = n_gs + n_start + n_sorted + al256(feat_t_bytes);
  if (total == 0) return RFD_OK;
  QgWorkspace w;
  const int rc = ws_get(st, use_grid ? B : 0, total, &w);
  if (rc != RFD_OK) return rc;
  uint8_t *p = w.scratch;
  if (use_grid) {
    g->gs = reinterpret_cast<GridScene *>(p); p += n_gs;
    g->start = reinterpret_cast<int *>(p); p += n_start;
    g->sorted = reinterpret_cast<float4 *>(p); p += n_sorted;
  }
  if (feat_t_bytes) g->feat_t = reinterpret_cast<float *>(p);
  if (use_grid) {
    RFD_CHECK_CUDA(cudaFuncSetAttribute(grid_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GRID_NC * 4),
                   "grid_build attr");
    grid_build_kernel<<<dim3(GB_CS, B), GB_THREADS, GRID_NC * 4, st>>>(xyz, N, radius, g->gs, w.count, g->start,
                                                                       g->sorted);
    RFD_CHECK_LAUNCH("grid_build_kernel");
  }
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
  cudaStream_t st = as_stream(stream);
  GridPtrs g;
  const bool use_grid = N >= GRID_MIN_N && radius > 0.f && nsample <= GRID_CAP / 2;
  const int rc = prepare(xyz, B, N, radius, use_grid, 0, st, &g);
  if (rc != RFD_OK) return rc;
  ball_query_kernel<<<grid, BQ_THREADS, 0, st>>>(new_xyz, xyz, N, M, radius, nsample, idx, g.gs, g.start, g.sorted);
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
  cudaStream_t st = as_stream(stream);
  QgParams P = {};
  P.xyz = xyz; P.new_xyz = new_xyz; P.features = features;
  P.n = N; P.m = M; P.C = C; P.nsample = nsample; P.radius = radius;
  P.use_xyz = use_xyz; P.normalize_xyz = normalize_xyz;
  P.new_features = new_features; P.grouped_xyz = grouped_xyz; P.idx_out = idx;
  P.s_shift = (nsample & (nsample - 1)) == 0 ? __builtin_ctz((unsigned)nsample) : -1;
  // queries per CTA: >= 8 (one per warp in phase 1), ~256 slots, never more than QG_MAX_SLOTS
  int qt = nsample >= 256 ? QG_MAX_SLOTS / nsample : 256 / nsample;
  if (nsample < 256 && qt < BQ_WARPS) qt = BQ_WARPS;
  if (qt < 1) qt = 1;
  if (qt > M) qt = M;
  P.qt = qt;
  const bool use_grid = N >= GRID_MIN_N && radius > 0.f;
  const bool transposed = C >= 8;  // below that the direct 4-byte gathers are cheaper than the extra pass
  P.Cp = transposed ? (C + 3) & ~3 : 0;
  GridPtrs g;
  const int rc = prepare(xyz, B, N, radius, use_grid, transposed ? sizeof(float) * (size_t)B * N * P.Cp : 0, st, &g);
  if (rc != RFD_OK) return rc;
  P.gs = g.gs; P.gstart = g.start; P.gsorted = g.sorted; P.feat_t = g.feat_t;
  if (transposed) {
    dim3 tg(h_ceil_div(N, 32), h_ceil_div(P.Cp, 32), B);
    if (tg.y > 65535) return RFD_ERR_UNSUPPORTED_SIZE;
    transpose_features_kernel<<<tg, 256, 0, st>>>(features, C, N, P.Cp, g.feat_t);
    RFD_CHECK_LAUNCH("transpose_features_kernel");
  }
  P.cloud_floats = use_grid ? 0 : 3 * (N < QG_CHUNK ? ((N + 3) & ~3) : QG_CHUNK);
  size_t smem = (size_t)qt * nsample * 4 + (size_t)qt * 3 * 4 + (size_t)qt * 8 + 16 + 16;
  smem += use_grid ? (size_t)BQ_WARPS * GRID_CAP * 4 : (size_t)P.cloud_floats * 4;
  smem += 128 + (transposed ? (size_t)BQ_WARPS * QG_TILE : 0);
  RFD_CHECK_CUDA(cudaFuncSetAttribute(query_and_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024),
                 "query_and_group attr");
  dim3 grid(h_ceil_div(M, qt), B);
  query_and_group_kernel<<<grid, BQ_THREADS, smem, st>>>(P);
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

extern "C" int rfd_transpose_features(const float *features, int B, int C, int N, int Cp, float *out, void *stream) {
  if (B < 0 || C < 0 || N < 0 || Cp < C || (Cp & 3)) return RFD_ERR_INVALID_ARGUMENT;
  if (B == 0 || N == 0 || Cp == 0) return RFD_OK;
  if (!features || !out) return RFD_ERR_INVALID_ARGUMENT;
  dim3 tg(h_ceil_div(N, 32), h_ceil_div(Cp, 32), B);
  if (tg.y > 65535 || B > 65535) return RFD_ERR_UNSUPPORTED_SIZE;
  transpose_features_kernel<<<tg, 256, 0, as_stream(stream)>>>(features, C, N, Cp, out);
  RFD_CHECK_LAUNCH("transpose_features_kernel");
  return RFD_OK;
}
