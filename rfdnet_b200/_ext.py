"""Drop-in replacement for the reference's pybind module `pointnet2_ops._ext`
(_ext-src/src/bindings.cpp:6-19): the same nine function names, argument order, return
types, zero-initialised outputs and precondition errors (utils.h:5-25 -> RuntimeError),
implemented on librfdnet_b200.so through the C ABI.  CUDA tensors only -- like the reference
("CPU not supported", e.g. ball_query.cpp:27-29) there is no CPU path.
"""
import torch

from . import _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _check(*specs):
    """specs: (tensor, name, dtype).  Same order as the reference wrappers (e.g. sampling.cpp:16-23):
    every CHECK_CONTIGUOUS, then every CHECK_IS_FLOAT/INT, then the device checks."""
    for t, name, _ in specs:
        if not isinstance(t, torch.Tensor):
            raise RuntimeError(f"{name} must be a tensor")
        if not t.is_contiguous():
            raise RuntimeError(f"{name} must be a contiguous tensor")
    for t, name, dtype in specs:
        if t.dtype != dtype:
            raise RuntimeError(f"{name} must be a{'n int' if dtype == torch.int32 else ' float'} tensor")
    first = specs[0][0]
    if not first.is_cuda:
        raise RuntimeError("CPU not supported")  # AT_ASSERT(false, "CPU not supported")
    for t, name, _ in specs[1:]:
        if t.device != first.device:
            raise RuntimeError(f"{name} must be a CUDA tensor")  # CHECK_CUDA


F32, I32 = torch.float32, torch.int32


def furthest_point_sampling(points, nsamples):
    """sampling.cpp:66-87.  points (B,N,3) f32 -> (B,nsamples) i32."""
    _check((points, "points", F32))
    B, N, _ = points.shape
    out = torch.empty((B, int(nsamples)), dtype=torch.int32, device=points.device)
    with torch.cuda.device(points.device):
        _lib.check(_lib.load().rfd_furthest_point_sampling(points.data_ptr(), B, N, int(nsamples), out.data_ptr(),
                                                            _stream()), "furthest_point_sampling")
    return out


def gather_points(points, idx):
    """sampling.cpp:15-38.  (B,C,N) f32, (B,M) i32 -> (B,C,M)."""
    _check((points, "points", F32), (idx, "idx", I32))
    B, C, N = points.shape
    M = idx.shape[1]
    out = torch.empty((B, C, M), dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        _lib.check(_lib.load().rfd_gather_points(points.data_ptr(), idx.data_ptr(), B, C, N, M, out.data_ptr(),
                                                  _stream()), "gather_points")
    return out


def gather_points_grad(grad_out, idx, n):
    """sampling.cpp:40-65.  (B,C,M), (B,M) -> (B,C,n)."""
    _check((grad_out, "grad_out", F32), (idx, "idx", I32))
    B, C, M = grad_out.shape
    out = torch.empty((B, C, int(n)), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        _lib.check(_lib.load().rfd_gather_points_grad(grad_out.data_ptr(), idx.data_ptr(), B, C, int(n), M,
                                                       out.data_ptr(), _stream()), "gather_points_grad")
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    """ball_query.cpp:8-32 (new_xyz FIRST).  (B,M,3), (B,N,3) -> (B,M,nsample) i32."""
    _check((new_xyz, "new_xyz", F32), (xyz, "xyz", F32))
    B, M, _ = new_xyz.shape
    N = xyz.shape[1]
    out = torch.empty((B, M, int(nsample)), dtype=torch.int32, device=new_xyz.device)
    with torch.cuda.device(new_xyz.device):
        _lib.check(_lib.load().rfd_ball_query(new_xyz.data_ptr(), xyz.data_ptr(), B, N, M, float(radius),
                                               int(nsample), out.data_ptr(), _stream()), "ball_query")
    return out


def group_points(points, idx):
    """group_points.cpp:12-36.  (B,C,N), (B,M,S) i32 -> (B,C,M,S)."""
    _check((points, "points", F32), (idx, "idx", I32))
    B, C, N = points.shape
    _, M, S = idx.shape
    out = torch.empty((B, C, M, S), dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        _lib.check(_lib.load().rfd_group_points(points.data_ptr(), idx.data_ptr(), B, C, N, M, S, out.data_ptr(),
                                                 _stream()), "group_points")
    return out


def group_points_grad(grad_out, idx, n):
    """group_points.cpp:38-62.  (B,C,M,S), (B,M,S) -> (B,C,n)."""
    _check((grad_out, "grad_out", F32), (idx, "idx", I32))
    B, C, M, S = grad_out.shape
    out = torch.empty((B, C, int(n)), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        _lib.check(_lib.load().rfd_group_points_grad(grad_out.data_ptr(), idx.data_ptr(), B, C, int(n), M, S,
                                                      out.data_ptr(), _stream()), "group_points_grad")
    return out


def three_nn(unknowns, knows):
    """interpolate.cpp:14-40.  (B,n,3), (B,m,3) -> [dist2 (B,n,3) f32, idx (B,n,3) i32]."""
    _check((unknowns, "unknowns", F32), (knows, "knows", F32))
    B, n, _ = unknowns.shape
    m = knows.shape[1]
    dist2 = torch.empty((B, n, 3), dtype=torch.float32, device=unknowns.device)
    idx = torch.empty((B, n, 3), dtype=torch.int32, device=unknowns.device)
    with torch.cuda.device(unknowns.device):
        _lib.check(_lib.load().rfd_three_nn(unknowns.data_ptr(), knows.data_ptr(), B, n, m, dist2.data_ptr(),
                                             idx.data_ptr(), _stream()), "three_nn")
    return [dist2, idx]


def three_interpolate(points, idx, weight):
    """interpolate.cpp:42-70.  (B,C,m), (B,n,3) i32, (B,n,3) f32 -> (B,C,n)."""
    _check((points, "points", F32), (idx, "idx", I32), (weight, "weight", F32))
    B, C, m = points.shape
    n = idx.shape[1]
    out = torch.empty((B, C, n), dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        _lib.check(_lib.load().rfd_three_interpolate(points.data_ptr(), idx.data_ptr(), weight.data_ptr(), B, C, m,
                                                      n, out.data_ptr(), _stream()), "three_interpolate")
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    """interpolate.cpp:71-100.  (B,C,n), (B,n,3), (B,n,3) -> (B,C,m)."""
    _check((grad_out, "grad_out", F32), (idx, "idx", I32), (weight, "weight", F32))
    B, C, n = grad_out.shape
    out = torch.empty((B, C, int(m)), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        _lib.check(_lib.load().rfd_three_interpolate_grad(grad_out.data_ptr(), idx.data_ptr(), weight.data_ptr(), B,
                                                           C, n, int(m), out.data_ptr(), _stream()),
                   "three_interpolate_grad")
    return out
