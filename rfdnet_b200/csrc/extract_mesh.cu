// extract_mesh.cu -- dense occupancy lattice -> triangle mesh on the GPU.
//
// Reference: models/iscnet/modules/generator.py:145-168 (Generator3D.extract_mesh), per object on the HOST:
//     occ_hat = logits.cpu().numpy().reshape(R,R,R)                  (131 KB D2H per object + a sync, :96-97,137-141)
//     occ_hat_padded = np.pad(occ_hat, 1, constant_values=-1e6)      (watertight border)
//     vertices, triangles = mcubes.marching_cubes(occ_hat_padded, threshold)      (PyMCubes 0.1.2, environment.yml:77)
//     vertices -= 0.5; vertices -= 1; vertices /= (R-1); vertices = box_size * (vertices - 0.5)
// Here ONE kernel does all of it for a batch of objects straight from the decoder's logits in HBM, so only the meshes
// (typically a fraction of the 32-bit lattice) ever cross PCIe.
//
// One CTA (1024 threads) per object; the padded lattice P = R + 2 (<= 34) lives in shared memory as
//   sign[p]  1 byte : value(p) <= threshold  (padding: -1e6 <= threshold)
//   info[p]  4 bytes: index of the first vertex owned by lattice point p (20 bits) | mask of its active +x,+y,+z edges
// A vertex belongs to the lattice edge it lies on, an edge to its lower end point ("owner") and axis.  Passes:
//   1  signs; 2  per point: active-edge mask, per cell: triangle count (thread-contiguous chunks) -> CTA-wide exclusive
//   scan -> info[]; 3  the object's vertex / triangle ranges are reserved in the output pools with one atomicAdd each;
//   4  vertices: interpolation in fp64 with the exact operation order of PyMCubes and of the numpy post-transform (no
//   FMA contraction), stored as f32 or f64; 5  triangles: Bourke's table per cell, vertex ids through info[].
// Output order (deterministic per object): vertices by (owner point, axis), triangles by cell then table order --
// PyMCubes' own triangle order; its vertex order differs (creation order), which trimesh(process=False) does not care
// about.  Object ranges inside the pools are handed out by atomics (any order), recorded in `ranges`.
#include <mutex>

#include "common.cuh"
#include "mc_tables.h"

namespace rfd {

constexpr int MC_THREADS = 1024;
constexpr int MC_MAX_P = 34;

__constant__ signed char c_tri_table[256][16];
__constant__ unsigned char c_tri_count[256];
// edge -> owner offset (di,dj,dk) and axis
__constant__ unsigned char c_edge_owner[12][4] = {{0, 0, 0, 0}, {1, 0, 0, 1}, {0, 1, 0, 0}, {0, 0, 0, 1}, {0, 0, 1, 0}, {1, 0, 1, 1},
                                                  {0, 1, 1, 0}, {0, 0, 1, 1}, {0, 0, 0, 2}, {1, 0, 0, 2}, {1, 1, 0, 2}, {0, 1, 0, 2}};

struct McParams {
  const float *logits;   // (B, R^3), x slowest, z fastest
  int B, R;
  double iso;            // threshold as the reference passes it to mcubes (python float)
  double box_size;
  void *vertices;        // pool: cap_v x 3 (f32 or f64)
  int vertex_f64;
  int *triangles;        // pool: cap_t x 3, object-local vertex ids
  long long cap_v, cap_t;
  int *ranges;           // (B, 4): vertex offset, vertex count, triangle offset, triangle count
  unsigned long long *totals;  // [0] vertices reserved, [1] triangles reserved, [2] objects that did not fit
};

__device__ __forceinline__ float mc_value(const float *__restrict__ lg, int R, int i, int j, int k) {
  // padded lattice point (i,j,k), 0 <= i,j,k < R+2
  if (i < 1 || j < 1 || k < 1 || i > R || j > R || k > R) return -1e6f;
  return __ldg(lg + ((size_t)(i - 1) * R + (j - 1)) * R + (k - 1));
}

// (x2-x1)*(iso-f1)/(f2-f1)+x1 -- mcubes/src/marchingcubes.h mc_isovalue_interpolation, double, no contraction
__device__ __forceinline__ double mc_interp(double iso, double f1, double f2, double x1, double x2) {
  if (f2 == f1) return __ddiv_rn(__dadd_rn(x2, x1), 2.0);
  return __dadd_rn(__ddiv_rn(__dmul_rn(__dsub_rn(x2, x1), __dsub_rn(iso, f1)), __dsub_rn(f2, f1)), x1);
}

// generator.py:162-168 on one coordinate: ((c - 0.5) - 1) / (R-1), then box_size * (. - 0.5)
__device__ __forceinline__ double mc_to_box(double c, double rm1, double box) {
  c = __dsub_rn(c, 0.5);
  c = __dsub_rn(c, 1.0);
  c = __ddiv_rn(c, rm1);
  return __dmul_rn(box, __dsub_rn(c, 0.5));
}

__global__ void __launch_bounds__(MC_THREADS, 1) extract_mesh_kernel(const McParams P) {
  extern __shared__ __align__(16) unsigned char mc_smem[];
  const int R = P.R, Pd = R + 2, PP = Pd * Pd, NP = PP * Pd;
  unsigned int *info = reinterpret_cast<unsigned int *>(mc_smem);            // NP
  unsigned char *sign = mc_smem + (size_t)NP * 4;                           // NP
  // the case tables are indexed by per-thread data: from constant memory every distinct address of a warp is a separate
  // (serialised) fetch -- 0.5 ms per 256 objects in the first version of this kernel; shared memory serves them at once
  __shared__ signed char s_tri_table[256][16];
  __shared__ unsigned char s_tri_count[256];
  __shared__ __align__(4) unsigned char s_edge_owner[12][4];
  for (int e = threadIdx.x; e < 256 * 16; e += MC_THREADS) (&s_tri_table[0][0])[e] = (&c_tri_table[0][0])[e];
  if (threadIdx.x < 256) s_tri_count[threadIdx.x] = c_tri_count[threadIdx.x];
  if (threadIdx.x < 48) (&s_edge_owner[0][0])[threadIdx.x] = (&c_edge_owner[0][0])[threadIdx.x];
  __shared__ int s_scan_v[32], s_scan_t[32];
  __shared__ long long s_off[2];
  __shared__ int s_tot[2], s_fit;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  const float *lg = P.logits + (size_t)b * R * R * R;
  // ---- 1: signs
  for (int p = tid; p < NP; p += MC_THREADS) {
    const int i = p / PP, r = p - i * PP, j = r / Pd, k = r - j * Pd;
    sign[p] = (double)mc_value(lg, R, i, j, k) <= P.iso ? 1 : 0;  // compared as mcubes does: the f32 value as double
  }
  __syncthreads();
  // ---- 2: counts over this thread's contiguous chunk of lattice points
  const int chunk = (NP + MC_THREADS - 1) / MC_THREADS;
  const int p0 = min(tid * chunk, NP), p1 = min(p0 + chunk, NP);
  auto edge_mask = [&](int p, int i, int j, int k) -> unsigned {
    const unsigned s = sign[p];
    unsigned m = 0;
    if (i + 1 < Pd && sign[p + PP] != s) m |= 1u;
    if (j + 1 < Pd && sign[p + Pd] != s) m |= 2u;
    if (k + 1 < Pd && sign[p + 1] != s) m |= 4u;
    return m;
  };
  auto cube_index = [&](int p) -> unsigned {
    return (unsigned)sign[p] | ((unsigned)sign[p + PP] << 1) | ((unsigned)sign[p + PP + Pd] << 2) | ((unsigned)sign[p + Pd] << 3) |
           ((unsigned)sign[p + 1] << 4) | ((unsigned)sign[p + PP + 1] << 5) | ((unsigned)sign[p + PP + Pd + 1] << 6) |
           ((unsigned)sign[p + Pd + 1] << 7);
  };
  int cv = 0, ct = 0;
  {
    int i = p0 / PP, r = p0 - i * PP, j = r / Pd, k = r - j * Pd;
    for (int p = p0; p < p1; ++p) {
      cv += __popc(edge_mask(p, i, j, k));
      if (i + 1 < Pd && j + 1 < Pd && k + 1 < Pd) ct += s_tri_count[cube_index(p)];
      if (++k == Pd) { k = 0; if (++j == Pd) { j = 0; ++i; } }
    }
  }
  // CTA-wide exclusive scan of (cv, ct)
  int iv = cv, it = ct;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int a = __shfl_up_sync(0xffffffffu, iv, off), c = __shfl_up_sync(0xffffffffu, it, off);
    if (lane >= off) { iv += a; it += c; }
  }
  if (lane == 31) { s_scan_v[warp] = iv; s_scan_t[warp] = it; }
  __syncthreads();
  if (warp == 0) {
    const int wv = s_scan_v[lane], wt = s_scan_t[lane];
    int a = wv, c = wt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int x = __shfl_up_sync(0xffffffffu, a, off), y = __shfl_up_sync(0xffffffffu, c, off);
      if (lane >= off) { a += x; c += y; }
    }
    s_scan_v[lane] = a - wv;
    s_scan_t[lane] = c - wt;
    if (lane == 31) { s_tot[0] = a; s_tot[1] = c; }
  }
  __syncthreads();
  int ov = s_scan_v[warp] + iv - cv;   // first vertex id of this thread's chunk
  const int ot = s_scan_t[warp] + it - ct;
  {
    int i = p0 / PP, r = p0 - i * PP, j = r / Pd, k = r - j * Pd;
    for (int p = p0; p < p1; ++p) {
      const unsigned m = edge_mask(p, i, j, k);
      info[p] = (unsigned)ov | (m << 20);
      ov += __popc(m);
      if (++k == Pd) { k = 0; if (++j == Pd) { j = 0; ++i; } }
    }
  }
  // ---- 3: reserve the object's ranges in the pools
  if (tid == 0) {
    const int nv = s_tot[0], nt = s_tot[1];
    const long long vo = (long long)atomicAdd(P.totals, (unsigned long long)nv);
    const long long to = (long long)atomicAdd(P.totals + 1, (unsigned long long)nt);
    const bool fit = vo + nv <= P.cap_v && to + nt <= P.cap_t && vo < 0x7fffffffLL && to < 0x7fffffffLL;
    if (!fit) atomicAdd(P.totals + 2, 1ull);
    s_off[0] = vo; s_off[1] = to; s_fit = fit ? 1 : 0;
    int *rg = P.ranges + (size_t)b * 4;
    rg[0] = fit ? (int)vo : -1; rg[1] = nv; rg[2] = fit ? (int)to : -1; rg[3] = nt;
  }
  __syncthreads();
  if (!s_fit) return;
  const long long voff = s_off[0], toff = s_off[1];
  // ---- 4: vertices
  {
    const double rm1 = (double)(R - 1);
    int i = p0 / PP, r = p0 - i * PP, j = r / Pd, k = r - j * Pd;
    for (int p = p0; p < p1; ++p) {
      const unsigned w = info[p], m = w >> 20;
      if (m) {
        long long vi = voff + (w & 0xfffffu);
        const double f0 = (double)mc_value(lg, R, i, j, k);
        const double base[3] = {(double)i, (double)j, (double)k};
#pragma unroll
        for (int ax = 0; ax < 3; ++ax) {
          if (!(m & (1u << ax))) continue;
          const double f1 = (double)mc_value(lg, R, i + (ax == 0), j + (ax == 1), k + (ax == 2));
          // PyMCubes creates an x-edge vertex as edge 6 of a cell (corner 6 -> 7: from the HIGHER x end), y / z edge
          // vertices as edges 5 / 10 (from the lower end)
          const double c = ax == 0 ? mc_interp(P.iso, f1, f0, base[0] + 1.0, base[0])
                                   : mc_interp(P.iso, f0, f1, base[ax], base[ax] + 1.0);
          double v[3];
#pragma unroll
          for (int d = 0; d < 3; ++d) v[d] = mc_to_box(d == ax ? c : base[d], rm1, P.box_size);
          if (P.vertex_f64) {
            double *o = reinterpret_cast<double *>(P.vertices) + vi * 3;
            o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
          } else {
            float *o = reinterpret_cast<float *>(P.vertices) + vi * 3;
            o[0] = (float)v[0]; o[1] = (float)v[1]; o[2] = (float)v[2];
          }
          ++vi;
        }
      }
      if (++k == Pd) { k = 0; if (++j == Pd) { j = 0; ++i; } }
    }
  }
  // ---- 5: triangles
  {
    int *tri = P.triangles + (toff + ot) * 3;
    int i = p0 / PP, r = p0 - i * PP, j = r / Pd, k = r - j * Pd;
    for (int p = p0; p < p1; ++p) {
      if (i + 1 < Pd && j + 1 < Pd && k + 1 < Pd) {
        const unsigned ci = cube_index(p);
        const int n = s_tri_count[ci];
        for (int t = 0; t < 3 * n; ++t) {
          const int e = s_tri_table[ci][t];
          const uchar4 eo = *reinterpret_cast<const uchar4 *>(s_edge_owner[e]);
          const int q = p + eo.x * PP + eo.y * Pd + eo.z;
          const unsigned w = info[q];
          *tri++ = (int)((w & 0xfffffu) + __popc((w >> 20) & ((1u << eo.w) - 1u)));
        }
      }
      if (++k == Pd) { k = 0; if (++j == Pd) { j = 0; ++i; } }
    }
  }
}

static std::once_flag g_mc_once[64];

static int mc_upload_tables() {
  unsigned char cnt[256];
  for (int c = 0; c < 256; ++c) {
    int n = 0;
    while (n < 16 && MC_TRI_TABLE[c][n] >= 0) ++n;
    cnt[c] = (unsigned char)(n / 3);
  }
  RFD_CHECK_CUDA(cudaMemcpyToSymbol(c_tri_table, MC_TRI_TABLE, sizeof(MC_TRI_TABLE)), "extract_mesh tables");
  RFD_CHECK_CUDA(cudaMemcpyToSymbol(c_tri_count, cnt, sizeof(cnt)), "extract_mesh tables");
  return RFD_OK;
}

}  // namespace rfd

using namespace rfd;

extern "C" int rfd_extract_mesh(const float *logits, int B, int R, double threshold, double box_size, void *vertices,
                                int vertex_f64, int *triangles, long long cap_vertices, long long cap_triangles,
                                int *ranges, unsigned long long *totals, void *stream) {
  if (B < 0 || R < 1 || cap_vertices < 0 || cap_triangles < 0) return RFD_ERR_INVALID_ARGUMENT;
  if (B == 0) return RFD_OK;
  if (!logits || !vertices || !triangles || !ranges || !totals) return RFD_ERR_INVALID_ARGUMENT;
  if (R + 2 > MC_MAX_P) return RFD_ERR_UNSUPPORTED_SIZE;  // the padded lattice must fit in shared memory
  int dev = 0;
  RFD_CHECK_CUDA(cudaGetDevice(&dev), "extract_mesh getdevice");
  if (dev < 0 || dev >= 64) return RFD_ERR_UNSUPPORTED_SIZE;
  int rc = RFD_OK;
  std::call_once(g_mc_once[dev], [&] { rc = mc_upload_tables(); });
  if (rc != RFD_OK) return rc;
  McParams P = {};
  P.logits = logits; P.B = B; P.R = R;
  P.iso = threshold; P.box_size = box_size;
  P.vertices = vertices; P.vertex_f64 = vertex_f64; P.triangles = triangles;
  P.cap_v = cap_vertices; P.cap_t = cap_triangles; P.ranges = ranges; P.totals = totals;
  const int np = (R + 2) * (R + 2) * (R + 2);
  const size_t smem = (size_t)np * 5 + 16;  // + 4.4 KB of static shared memory (tables)
  RFD_CHECK_CUDA(cudaFuncSetAttribute(extract_mesh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                 "extract_mesh attr");
  extract_mesh_kernel<<<B, MC_THREADS, smem, as_stream(stream)>>>(P);
  RFD_CHECK_LAUNCH("extract_mesh_kernel");
  return RFD_OK;
}
