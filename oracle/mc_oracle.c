/* mc_oracle.c -- CPU restatement of the surface-extraction step of the reference.
 * TEST INFRASTRUCTURE ONLY: linked into oracle/liboracle.so; the product (rfdnet_b200) never calls it.
 *
 * Reference call site: models/iscnet/modules/generator.py:145-168 (Generator3D.extract_mesh):
 *     occ_hat_padded = np.pad(occ_hat, 1, 'constant', constant_values=-1e6)
 *     vertices, triangles = mcubes.marching_cubes(occ_hat_padded, threshold)
 * `mcubes` is the third-party package PyMCubes, pinned by the reference at pymcubes==0.1.2 (environment.yml:77); it is
 * NOT vendored under /root/reference and not installed in this image, so its algorithm is restated here from the
 * published source (mcubes/src/marchingcubes.h, marching_cubes<>()):
 *   - the volume is sampled at integer lattice positions (lower = 0, upper = n-1 => dx = 1), values as double;
 *   - cells are visited with i (x) slowest, k (z) fastest; corner m of a cell has its bit set iff v[m] <= isovalue;
 *   - a cell creates the vertices of its edges 6, 5 and 10 (the three edges meeting at its corner 6), all other edges
 *     are looked up in the shared-index planes filled by earlier cells (or created on the i/j/k == 0 boundary);
 *   - a vertex lies at (x2-x1)*(iso-f1)/(f2-f1)+x1 along its edge, evaluated in double; edge 6 is walked from corner 6
 *     to corner 7 (decreasing x), edges 5 and 10 from the lower to the higher coordinate;
 *   - triangles are emitted per cell in the order of Bourke's table.
 * PARITY UNPINNED against the real PyMCubes binary (absent here); what is pinned: the table itself (edge sets, closed +
 * consistently oriented surfaces on random fields, tests/test_mesh_cpu.py) and hand-made known answers.
 * Besides PyMCubes' own output order this oracle reports, per vertex, the key 3*owner_point + axis of the lattice edge
 * it lies on, so that tests can compare against any other deterministic vertex order. */
#include "mc_tables_oracle.h"
#include <stdlib.h>

static const int MC_CORNER[8][3] = {{0,0,0},{1,0,0},{1,1,0},{0,1,0},{0,0,1},{1,0,1},{1,1,1},{0,1,1}};
/* edge -> (first corner, second corner) in the direction PyMCubes interpolates when IT CREATES the vertex */
static const int MC_EDGE_DIR[12][2] = {{0,1},{1,2},{2,3},{3,0},{4,5},{5,6},{6,7},{7,4},{0,4},{1,5},{2,6},{3,7}};

int oracle_marching_cubes(const double *vol, int nx, int ny, int nz, double iso, double *verts, int *keys, int cap_v,
                          int *tris, int cap_t, int *nv_out, int *nt_out) {
  const long long npts = (long long)nx * ny * nz;
  int *map = (int *)malloc(sizeof(int) * 3 * (size_t)npts);
  if (!map) return -1;
  for (long long e = 0; e < 3 * npts; ++e) map[e] = -1;
  int nv = 0, nt = 0, rc = 0;
  static const int order[12] = {6, 5, 10, 0, 1, 2, 3, 4, 7, 8, 9, 11};
  for (int i = 0; i < nx - 1; ++i)
    for (int j = 0; j < ny - 1; ++j)
      for (int k = 0; k < nz - 1; ++k) {
        double v[8];
        long long pt[8];
        int ci = 0;
        for (int m = 0; m < 8; ++m) {
          pt[m] = ((long long)(i + MC_CORNER[m][0]) * ny + (j + MC_CORNER[m][1])) * nz + (k + MC_CORNER[m][2]);
          v[m] = vol[pt[m]];
          if (v[m] <= iso) ci |= 1 << m;
        }
        const signed char *row = MC_TRI_TABLE[ci];
        int mask = 0;
        for (int t = 0; t < 16 && row[t] >= 0; ++t) mask |= 1 << row[t];
        int index[12];
        for (int o = 0; o < 12; ++o) {
          const int e = order[o];
          if (!(mask & (1 << e))) continue;
          const int a = MC_EDGE_DIR[e][0], b = MC_EDGE_DIR[e][1];
          int axis = 0;
          for (int d = 0; d < 3; ++d) if (MC_CORNER[a][d] != MC_CORNER[b][d]) axis = d;
          const long long owner = pt[a] < pt[b] ? pt[a] : pt[b];
          const long long key = owner * 3 + axis;
          if (map[key] < 0) {
            if (nv >= cap_v) { rc = 1; goto done; }
            const double pa[3] = {(double)(i + MC_CORNER[a][0]), (double)(j + MC_CORNER[a][1]), (double)(k + MC_CORNER[a][2])};
            const double x1 = pa[axis], x2 = (double)((axis == 0 ? i : axis == 1 ? j : k) + MC_CORNER[b][axis]);
            const double f1 = v[a], f2 = v[b];
            double c = (f2 == f1) ? (x2 + x1) / 2 : (x2 - x1) * (iso - f1) / (f2 - f1) + x1;
            verts[3 * nv + 0] = pa[0]; verts[3 * nv + 1] = pa[1]; verts[3 * nv + 2] = pa[2];
            verts[3 * nv + axis] = c;
            keys[nv] = (int)key;
            map[key] = nv++;
          }
          index[e] = map[key];
        }
        for (int t = 0; t < 16 && row[t] >= 0; t += 3) {
          if (nt >= cap_t) { rc = 1; goto done; }
          tris[3 * nt + 0] = index[row[t]]; tris[3 * nt + 1] = index[row[t + 1]]; tris[3 * nt + 2] = index[row[t + 2]];
          ++nt;
        }
      }
done:
  free(map);
  *nv_out = nv; *nt_out = nt;
  return rc;
}

const signed char *oracle_mc_table(void) { return &MC_TRI_TABLE[0][0]; }
