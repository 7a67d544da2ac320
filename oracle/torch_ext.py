"""`pointnet2_ops._ext`-shaped adapter over the CPU oracle for torch CPU tensors
(bindings.cpp:6-19 names).  TEST INFRASTRUCTURE ONLY: lets the UNMODIFIED reference Python
modules run on CPU (golden generation) and is the CPU baseline of bench.py."""
import numpy as np
import torch

from . import cpu_ref as R


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def furthest_point_sampling(points, nsamples):
    return _t(R.furthest_point_sampling(points.detach().numpy(), int(nsamples)))


def gather_points(points, idx):
    return _t(R.gather_points(points.detach().numpy(), idx.numpy()))


def gather_points_grad(grad_out, idx, n):
    return _t(R.gather_points_grad(grad_out.detach().numpy(), idx.numpy(), int(n)))


def ball_query(new_xyz, xyz, radius, nsample):
    return _t(R.ball_query(new_xyz.detach().numpy(), xyz.detach().numpy(), float(radius), int(nsample)))


def group_points(points, idx):
    return _t(R.group_points(points.detach().numpy(), idx.numpy()))


def group_points_grad(grad_out, idx, n):
    return _t(R.group_points_grad(grad_out.detach().numpy(), idx.numpy(), int(n)))


def three_nn(unknowns, knows):
    d2, ix = R.three_nn(unknowns.detach().numpy(), knows.detach().numpy())
    return [_t(d2), _t(ix)]


def three_interpolate(points, idx, weight):
    return _t(R.three_interpolate(points.detach().numpy(), idx.numpy(), weight.detach().numpy()))


def three_interpolate_grad(grad_out, idx, weight, m):
    return _t(R.three_interpolate_grad(grad_out.detach().numpy(), idx.numpy(), weight.detach().numpy(), int(m)))
