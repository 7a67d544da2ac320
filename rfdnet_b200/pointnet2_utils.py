"""Host-side mirror of the reference's `pointnet2_ops.pointnet2_utils`
(external/pointnet2_ops_lib/pointnet2_ops/pointnet2_utils.py:34-411): the same autograd
Functions and grouper modules, same names / argument order / return types, running on
rfdnet_b200._ext (hand-written sm_100a kernels behind the C ABI)."""
import torch
import torch.nn as nn
from torch.autograd import Function

from . import _ext, _lib


def _operator(name, ref_lines, forward, backward=None):
    """A torch.autograd.Function named like the reference's (pointnet2_utils.py:<ref_lines>) around one `_ext` entry
    point.  forward(ctx, *args) -> outputs; index-valued operators (backward=None) mark every output non-differentiable,
    the others pass backward(ctx, *grads) -> input gradients, computed by the scatter-add kernels of the library."""

    def fwd(ctx, *args):
        out = forward(ctx, *args)
        if backward is None:
            ctx.mark_non_differentiable(*(out if isinstance(out, tuple) else (out,)))
        return out

    def bwd(ctx, *grads):
        return () if backward is None else backward(ctx, *grads)

    return type(name, (Function,), {"forward": staticmethod(fwd), "backward": staticmethod(bwd),
                                    "__doc__": f"pointnet2_utils.py:{ref_lines}"})


def _saving(op, *keep):
    """forward that stashes the named positional inputs for the backward pass before calling `op`"""
    def forward(ctx, *args):
        ctx.save_for_backward(*(args[i] for i in keep))
        return op(*args)
    return forward


def _three_nn(ctx, unknown, known):
    dist2, idx = _ext.three_nn(unknown, known)
    return torch.sqrt(dist2), idx                      # the reference returns distances, the kernel squared ones (:125)


def _gather_grad(ctx, g):
    idx, features = ctx.saved_tensors
    return _ext.gather_points_grad(g.contiguous(), idx, features.size(2)), None


def _group_grad(ctx, g):
    idx, features = ctx.saved_tensors
    return _ext.group_points_grad(g.contiguous(), idx, features.size(2)), torch.zeros_like(idx)


def _interpolate_grad(ctx, g):
    idx, weight, features = ctx.saved_tensors
    return (_ext.three_interpolate_grad(g.contiguous(), idx, weight, features.size(2)), torch.zeros_like(idx),
            torch.zeros_like(weight))


FurthestPointSampling = _operator("FurthestPointSampling", "34-65", lambda ctx, xyz, npoint: _ext.furthest_point_sampling(xyz, npoint))
GatherOperation = _operator("GatherOperation", "68-101", _saving(_ext.gather_points, 1, 0), _gather_grad)
ThreeNN = _operator("ThreeNN", "104-136", _three_nn)
ThreeInterpolate = _operator("ThreeInterpolate", "139-191", _saving(_ext.three_interpolate, 1, 2, 0), _interpolate_grad)
GroupingOperation = _operator("GroupingOperation", "194-240", _saving(_ext.group_points, 1, 0), _group_grad)
# the Python-side argument order of the reference is (radius, nsample, xyz, new_xyz); the extension takes new_xyz first
BallQuery = _operator("BallQuery", "243-276", lambda ctx, radius, nsample, xyz, new_xyz: _ext.ball_query(new_xyz, xyz, radius, nsample))

furthest_point_sample = FurthestPointSampling.apply
gather_operation = GatherOperation.apply
three_nn = ThreeNN.apply
three_interpolate = ThreeInterpolate.apply
grouping_operation = GroupingOperation.apply
ball_query = BallQuery.apply


def fps_with_xyz(xyz, npoint, try_prefix=False):
    """FPS that also returns the sampled coordinates (B,npoint,3) straight from the kernel (inference path).

    try_prefix: the caller expects `xyz` to be in FPS order already (SA(k+1) samples the points SA(k) sampled,
    pointnet2backbone.py:104-113), in which case the answer is 0..npoint-1.  That is PROVED per scene by
    rfd_fps_prefix_check (a parallel sweep with the sampler's own arithmetic and tie-break, ~10 us) and the serial sampler
    is skipped for the scenes where it holds; the others run it.  The result is identical either way."""
    _mlp_check_f32(xyz, "points")
    B, N, _ = xyz.shape
    npoint = int(npoint)
    idx = torch.empty((B, npoint), dtype=torch.int32, device=xyz.device)
    new_xyz = torch.empty((B, npoint, 3), dtype=torch.float32, device=xyz.device)
    lib = _lib.load()
    with torch.cuda.device(xyz.device), _lib.timed("fps", float(B) * (12.0 * N + 4.0 * npoint)):
        st = torch.cuda.current_stream().cuda_stream
        flag = 0
        if try_prefix and npoint <= N:
            ws = torch.empty((B, max(npoint, 1)), dtype=torch.float32, device=xyz.device)
            flag_t = torch.empty((B,), dtype=torch.int32, device=xyz.device)
            _lib.check(lib.rfd_fps_prefix_check(xyz.data_ptr(), B, N, npoint, ws.data_ptr(), flag_t.data_ptr(), st),
                       "fps_prefix_check")
            flag = flag_t.data_ptr()
        _lib.check(lib.rfd_furthest_point_sampling_cond(xyz.data_ptr(), B, N, npoint, flag, idx.data_ptr(),
                                                        new_xyz.data_ptr(), st), "fps")
    return idx, new_xyz


def _mlp_check_f32(t, name):
    from .mlp import check_f32
    check_f32(t, name)


def fused_query_and_group(xyz, new_xyz, features, radius, nsample, use_xyz=True, normalize_xyz=False,
                          ret_grouped_xyz=False, ret_idx=False):
    """One kernel for QueryAndGroup.forward (pointnet2_utils.py:319-344): ball query + both gathers +
    centring + radius normalisation + concatenation.  Inference path (no autograd graph)."""
    _mlp_check_f32(xyz, "xyz")
    _mlp_check_f32(new_xyz, "new_xyz")
    if features is not None:
        _mlp_check_f32(features, "features")
        if features.device != xyz.device:
            raise RuntimeError("features must be a CUDA tensor")
    if new_xyz.device != xyz.device:
        raise RuntimeError("new_xyz must be a CUDA tensor")
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    C = 0 if features is None else features.shape[1]
    Ct = (3 if use_xyz else 0) + C
    dev = xyz.device
    out = torch.empty((B, Ct, M, nsample), dtype=torch.float32, device=dev)
    gxyz = torch.empty((B, 3, M, nsample), dtype=torch.float32, device=dev) if ret_grouped_xyz else None
    idx = torch.empty((B, M, nsample), dtype=torch.int32, device=dev) if ret_idx else None
    # algorithmic HBM bytes: 12N + 12M + 4CN + 4(3+C)MS (+ 4MS when idx is written)
    work = B * (12 * N + 12 * M + 4 * C * N + 4 * Ct * M * nsample + (4 * M * nsample if ret_idx else 0))
    with torch.cuda.device(dev), _lib.timed("query_and_group", work):
        _lib.check(_lib.load().rfd_query_and_group(
            xyz.data_ptr(), new_xyz.data_ptr(), 0 if features is None else features.data_ptr(), B, N, M, C,
            float(radius), int(nsample), int(bool(use_xyz)), int(bool(normalize_xyz)), out.data_ptr(),
            0 if gxyz is None else gxyz.data_ptr(), 0 if idx is None else idx.data_ptr(),
            torch.cuda.current_stream().cuda_stream), "query_and_group")
    return out, gxyz, idx


def _resample_uniformly(idx):
    """`sample_uniformly` of the reference (pointnet2_utils.py:321-330): every row of idx (B,M,S) keeps its distinct
    indices (ascending) and fills the remaining slots with uniform random draws from them.  Whole-tensor form on the
    device (sort, first-occurrence mask, stable compaction, one randint) instead of the reference's B*M host loop.
    -> (idx (B,M,S) i32, unique_cnt (B,M) f32)."""
    B, M, S = idx.shape
    srt, _ = torch.sort(idx.long(), dim=-1)
    first = torch.ones_like(srt, dtype=torch.bool)
    first[..., 1:] = srt[..., 1:] != srt[..., :-1]
    cnt = first.sum(-1)                                                   # distinct indices per row
    # stable compaction: distinct values to the front, in ascending order
    order = torch.argsort((~first).to(torch.int8), dim=-1, stable=True)
    uniq = torch.gather(srt, -1, order)
    slot = torch.arange(S, device=idx.device).view(1, 1, S)
    draw = (torch.rand((B, M, S), device=idx.device) * cnt.unsqueeze(-1)).long().clamp_(max=S - 1)
    draw = torch.minimum(draw, (cnt - 1).unsqueeze(-1))
    out = torch.where(slot < cnt.unsqueeze(-1), uniq, torch.gather(uniq, -1, draw))
    return out.to(torch.int32), cnt.to(torch.float32)


class QueryAndGroup(nn.Module):
    """pointnet2_utils.py:279-361.  Under autograd (training) the reference's op sequence runs on the
    drop-in kernels so gradients flow exactly as in the reference; without grad the fused kernel is used."""

    def __init__(self, radius, nsample, use_xyz=True, ret_grouped_xyz=False, normalize_xyz=False,
                 sample_uniformly=False, ret_unique_cnt=False):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz
        self.ret_grouped_xyz = ret_grouped_xyz
        self.normalize_xyz = normalize_xyz
        self.sample_uniformly = sample_uniformly
        self.ret_unique_cnt = ret_unique_cnt
        if self.ret_unique_cnt:
            assert self.sample_uniformly

    def forward(self, xyz, new_xyz, features=None):
        needs_grad = torch.is_grad_enabled() and (
            xyz.requires_grad or new_xyz.requires_grad or (features is not None and features.requires_grad))
        if features is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
        if not needs_grad and not self.sample_uniformly and self.nsample <= 1024 and xyz.is_cuda:
            new_features, grouped_xyz, _ = fused_query_and_group(
                xyz.contiguous(), new_xyz.contiguous(), None if features is None else features.contiguous(),
                self.radius, self.nsample, self.use_xyz, self.normalize_xyz, ret_grouped_xyz=self.ret_grouped_xyz)
            return (new_features, grouped_xyz) if self.ret_grouped_xyz else new_features

        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        if self.sample_uniformly:
            idx, unique_cnt = _resample_uniformly(idx)
        xyz_trans = xyz.transpose(1, 2).contiguous()
        grouped_xyz = grouping_operation(xyz_trans, idx)
        grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if self.normalize_xyz:
            grouped_xyz = grouped_xyz / self.radius
        if features is not None:
            grouped_features = grouping_operation(features, idx)
            new_features = torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
        else:
            new_features = grouped_xyz
        ret = [new_features]
        if self.ret_grouped_xyz:
            ret.append(grouped_xyz)
        if self.ret_unique_cnt:
            ret.append(unique_cnt)
        return ret[0] if len(ret) == 1 else tuple(ret)


class GroupAll(nn.Module):
    """pointnet2_utils.py:364-411"""

    def __init__(self, use_xyz=True, ret_grouped_xyz=False):
        super().__init__()
        self.use_xyz = use_xyz
        self.ret_grouped_xyz = ret_grouped_xyz

    def forward(self, xyz, new_xyz, features=None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is not None:
            grouped_features = features.unsqueeze(2)
            new_features = torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
        else:
            new_features = grouped_xyz
        return (new_features, grouped_xyz) if self.ret_grouped_xyz else new_features
