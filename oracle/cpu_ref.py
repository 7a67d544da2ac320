"""ctypes front-end of oracle/liboracle.so with the reference's `_ext` signatures
(bindings.cpp:6-19), on numpy arrays.  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os

import numpy as np

from . import build as _build

_lib = None


def _load():
    global _lib
    if _lib is None:
        path = _build.build()
        _lib = ctypes.CDLL(path)
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def opt_n_threads(n):
    return int(_load().oracle_opt_n_threads(int(n)))


def num_threads():
    return int(_load().oracle_num_threads())


def furthest_point_sampling(points, nsamples):
    """sampling.cpp:66-87: points (B,N,3) f32 -> (B,nsamples) i32."""
    p, pp = _f(points)
    B, N, _ = p.shape
    out = np.zeros((B, nsamples), np.int32)
    rc = _load().oracle_furthest_point_sampling(pp, B, N, int(nsamples), out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return out


def gather_points(points, idx):
    """sampling.cpp:15-38: (B,C,N), (B,M) -> (B,C,M)."""
    p, pp = _f(points)
    ix, ip = _i(idx)
    B, C, N = p.shape
    M = ix.shape[1]
    out = np.zeros((B, C, M), np.float32)
    _load().oracle_gather_points(pp, ip, B, C, N, M, out.ctypes.data_as(ctypes.c_void_p))
    return out


def gather_points_grad(grad_out, idx, n):
    g, gp = _f(grad_out)
    ix, ip = _i(idx)
    B, C, M = g.shape
    out = np.zeros((B, C, n), np.float32)
    _load().oracle_gather_points_grad(gp, ip, B, C, int(n), M, out.ctypes.data_as(ctypes.c_void_p))
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    """ball_query.cpp:8-32 (note: new_xyz first)."""
    q, qp = _f(new_xyz)
    p, pp = _f(xyz)
    B, M, _ = q.shape
    N = p.shape[1]
    out = np.zeros((B, M, nsample), np.int32)
    _load().oracle_ball_query(qp, pp, B, N, M, ctypes.c_float(radius), int(nsample),
                              out.ctypes.data_as(ctypes.c_void_p))
    return out


def group_points(points, idx):
    """group_points.cpp:12-36: (B,C,N), (B,M,S) -> (B,C,M,S)."""
    p, pp = _f(points)
    ix, ip = _i(idx)
    B, C, N = p.shape
    _, M, S = ix.shape
    out = np.zeros((B, C, M, S), np.float32)
    _load().oracle_group_points(pp, ip, B, C, N, M, S, out.ctypes.data_as(ctypes.c_void_p))
    return out


def group_points_grad(grad_out, idx, n):
    g, gp = _f(grad_out)
    ix, ip = _i(idx)
    B, C, M, S = g.shape
    out = np.zeros((B, C, n), np.float32)
    _load().oracle_group_points_grad(gp, ip, B, C, int(n), M, S, out.ctypes.data_as(ctypes.c_void_p))
    return out


def three_nn(unknowns, knows):
    """interpolate.cpp:14-40: -> (dist2 (B,n,3) f32, idx (B,n,3) i32)."""
    u, up = _f(unknowns)
    k, kp = _f(knows)
    B, n, _ = u.shape
    m = k.shape[1]
    d2 = np.zeros((B, n, 3), np.float32)
    ix = np.zeros((B, n, 3), np.int32)
    _load().oracle_three_nn(up, kp, B, n, m, d2.ctypes.data_as(ctypes.c_void_p), ix.ctypes.data_as(ctypes.c_void_p))
    return d2, ix


def three_interpolate(points, idx, weight):
    """interpolate.cpp:42-70: (B,C,m), (B,n,3), (B,n,3) -> (B,C,n)."""
    p, pp = _f(points)
    ix, ip = _i(idx)
    w, wp = _f(weight)
    B, C, m = p.shape
    n = ix.shape[1]
    out = np.zeros((B, C, n), np.float32)
    _load().oracle_three_interpolate(pp, ip, wp, B, C, m, n, out.ctypes.data_as(ctypes.c_void_p))
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    g, gp = _f(grad_out)
    ix, ip = _i(idx)
    w, wp = _f(weight)
    B, C, n = g.shape
    out = np.zeros((B, C, m), np.float32)
    _load().oracle_three_interpolate_grad(gp, ip, wp, B, C, n, int(m), out.ctypes.data_as(ctypes.c_void_p))
    return out


def query_and_group(xyz, new_xyz, features, radius, nsample, use_xyz=True, normalize_xyz=False):
    """QueryAndGroup.forward, pointnet2_utils.py:302-361 (sample_uniformly=False).

    Returns (new_features (B,3+C,M,S), grouped_xyz (B,3,M,S), idx).  The in-place
    `grouped_xyz /= radius` (:337) is evaluated the way torch does for a Python-float
    divisor on an f32 tensor: true division by float32(radius) per element (IEEE div)."""
    xyz = np.ascontiguousarray(xyz, np.float32)
    new_xyz = np.ascontiguousarray(new_xyz, np.float32)
    idx = ball_query(new_xyz, xyz, radius, nsample)
    grouped_xyz = group_points(np.ascontiguousarray(xyz.transpose(0, 2, 1)), idx)
    grouped_xyz = grouped_xyz - new_xyz.transpose(0, 2, 1)[..., None]
    if normalize_xyz:
        grouped_xyz = (grouped_xyz / np.float32(radius)).astype(np.float32)
    if features is not None:
        gf = group_points(features, idx)
        new_features = np.concatenate([grouped_xyz, gf], axis=1) if use_xyz else gf
    else:
        new_features = grouped_xyz
    return new_features, grouped_xyz, idx


def marching_cubes(vol, iso):
    """mcubes.marching_cubes(vol, iso) restated (oracle/mc_oracle.c): vol (nx,ny,nz) -> vertices (V,3) f64 in lattice
    index coordinates, triangles (T,3) i32 into them, keys (V,) i32 = 3 * owner lattice point + axis of each vertex's edge."""
    v = np.ascontiguousarray(vol, dtype=np.float64)
    nx, ny, nz = v.shape
    cap_v, cap_t = 3 * v.size + 8, 5 * v.size + 8
    verts = np.zeros((cap_v, 3), np.float64)
    keys = np.zeros(cap_v, np.int32)
    tris = np.zeros((cap_t, 3), np.int32)
    nv, nt = ctypes.c_int(0), ctypes.c_int(0)
    lib = _load()
    lib.oracle_marching_cubes.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                          ctypes.c_void_p, ctypes.c_void_p]
    rc = lib.oracle_marching_cubes(v.ctypes.data_as(ctypes.c_void_p), nx, ny, nz, float(iso),
                                   verts.ctypes.data_as(ctypes.c_void_p), keys.ctypes.data_as(ctypes.c_void_p), cap_v,
                                   tris.ctypes.data_as(ctypes.c_void_p), cap_t, ctypes.byref(nv), ctypes.byref(nt))
    assert rc == 0
    return verts[:nv.value].copy(), tris[:nt.value].copy(), keys[:nv.value].copy()


def extract_mesh(occ_hat, threshold=0.5, padding=0.1):
    """Generator3D.extract_mesh (generator.py:145-168) up to the trimesh construction: occ_hat (n,n,n) logits ->
    vertices (V,3) f64 in the object's box frame, triangles (T,3), keys (V,) (see marching_cubes; padded lattice)."""
    occ_hat = np.asarray(occ_hat)
    n_x, n_y, n_z = occ_hat.shape
    box_size = 1 + padding
    thr = np.log(threshold) - np.log(1. - threshold)
    occ_hat_padded = np.pad(occ_hat, 1, 'constant', constant_values=-1e6)
    vertices, triangles, keys = marching_cubes(occ_hat_padded, thr)
    vertices -= 0.5
    vertices -= 1
    vertices /= np.array([n_x - 1, n_y - 1, n_z - 1])
    vertices = box_size * (vertices - 0.5)
    return vertices, triangles, keys


def mc_table():
    lib = _load()
    lib.oracle_mc_table.restype = ctypes.POINTER(ctypes.c_byte)
    p = lib.oracle_mc_table()
    return np.ctypeslib.as_array(p, shape=(256, 16)).astype(np.int8).copy()
