"""Host-side mirror of the mesh-extraction half of Generator3D (models/iscnet/modules/generator.py:64-168), batched and
on the device: decoder logits on the dense R^3 lattice -> triangle meshes, without the per-object `.cpu()` of the value
grid, `np.pad` and `mcubes.marching_cubes` of the reference (generator.py:96-97,137-141,157-161).

  MeshBatch            vertices / triangles pools on the device + per-object ranges; `.to_host()` copies only what is used
  extract_meshes       logits (B, R^3) -> MeshBatch           (rfd_extract_mesh, csrc/extract_mesh.cu)
  logit_threshold      generator.py:85: log(t) - log(1 - t)

Out of scope here: MISE up-sampling (upsampling_steps > 0), mesh simplification, refinement and
normal estimation -- the reference's ISCNet config runs upsampling_steps = 0 and none of the latter.
"""
import math

import torch

from . import _lib


def logit_threshold(threshold=0.5):
    return math.log(threshold) - math.log(1.0 - threshold)


class MeshBatch:
    """Meshes of B objects in two device pools.  ranges (B,4) i32: vertex offset, vertex count, triangle offset, triangle
    count (offset -1: the object did not fit -> `overflowed`); triangles hold object-local vertex ids."""

    def __init__(self, vertices, triangles, ranges, totals):
        self.vertices, self.triangles, self.ranges, self.totals = vertices, triangles, ranges, totals
        self._host = None

    def to_host(self, stream_sync=True):
        """-> (vertices (V,3) numpy, triangles (T,3) numpy, ranges (B,4) numpy); copies the used prefix of the pools only.
        One small D2H (ranges + totals) decides how much of the pools is copied."""
        if self._host is None:
            tot = self.totals.cpu()
            rng = self.ranges.cpu().numpy()
            if int(tot[2]) != 0:
                raise RuntimeError(f"extract_meshes: {int(tot[2])} object(s) did not fit in the pools "
                                   f"(needed {int(tot[0])} vertices / {int(tot[1])} triangles); pass larger capacities")
            nv, nt = int(tot[0]), int(tot[1])
            self._host = (self.vertices[:nv].cpu().numpy(), self.triangles[:nt].cpu().numpy(), rng)
        return self._host

    def d2h_bytes(self):
        v, t, r = self.to_host()
        return v.nbytes + t.nbytes + r.nbytes + 24

    def mesh(self, b):
        """(vertices (n,3), triangles (m,3)) of object b -- what trimesh.Trimesh(vertices, triangles, process=False) takes."""
        v, t, r = self.to_host()
        vo, nv, to, nt = (int(x) for x in r[b])
        return v[vo:vo + nv], t[to:to + nt]

    def __len__(self):
        return self.ranges.shape[0]


def extract_meshes(logits, resolution=32, threshold=0.5, padding=0.1, vertex_dtype=torch.float32,
                   vertices_per_object=8192, triangles_per_object=16384, pools=None, ranges=None, totals=None):
    """Generator3D.extract_mesh for every row of logits (B, R^3) f32 cuda (x slowest, z fastest).
    `pools` = (vertices, triangles) preallocated device tensors to reuse between calls; `ranges` (B,4) i32 / `totals`
    (3,) i64 let several calls (object chunks of one step) append to the same pools -- `totals` is zeroed by whoever
    starts the step."""
    if not logits.is_cuda:
        raise RuntimeError("CPU not supported")
    if logits.dtype != torch.float32:
        raise RuntimeError("logits must be a float tensor")
    logits = logits.contiguous()
    B, T = logits.shape
    R = int(resolution)
    if T != R ** 3:
        raise ValueError(f"logits has {T} values per object, expected {R}^3")
    dev = logits.device
    f64 = vertex_dtype == torch.float64
    if pools is None:
        vertices = torch.empty((max(1, B * vertices_per_object), 3), dtype=vertex_dtype, device=dev)
        triangles = torch.empty((max(1, B * triangles_per_object), 3), dtype=torch.int32, device=dev)
    else:
        vertices, triangles = pools
        assert vertices.dtype == vertex_dtype and triangles.dtype == torch.int32 and vertices.is_contiguous()
    if ranges is None:
        ranges = torch.empty((B, 4), dtype=torch.int32, device=dev)
    if totals is None:
        totals = torch.zeros((3,), dtype=torch.int64, device=dev)
    assert ranges.shape == (B, 4) and ranges.dtype == torch.int32 and ranges.is_contiguous()
    with torch.cuda.device(dev):
        _lib.check(_lib.load().rfd_extract_mesh(
            logits.data_ptr(), B, R, float(logit_threshold(threshold)), float(1 + padding), vertices.data_ptr(), int(f64),
            triangles.data_ptr(), vertices.shape[0], triangles.shape[0], ranges.data_ptr(), totals.data_ptr(),
            torch.cuda.current_stream().cuda_stream), "extract_mesh")
    return MeshBatch(vertices, triangles, ranges, totals)
