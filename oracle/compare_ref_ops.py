"""Timing of rfdnet_b200 against the UNMODIFIED reference pointnet2_ops CUDA kernels recompiled for sm_100
(oracle/_ref/_ref_ext.so) on identical inputs, same GPU: the point-cloud front end of every SA layer
(FPS -> gather -> ball query -> group xyz -> sub -> div -> group features -> cat), i.e. pointnet2_modules.py:219-229 +
pointnet2_utils.py:319-344, driven exactly as the reference's Python does.  Results are checked equal before timing.
Measurement infrastructure only (imports oracle/)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import build_ref
from rfdnet_b200 import _ext, pointnet2_utils
from rfdnet_b200.synth import scannet_like_batch

ref = build_ref.load()
assert ref is not None, "oracle/_ref/_ref_ext.so missing"
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4


def timeit(fn, iters=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, out


def ref_front(xyz, feats, npoint, radius, nsample):
    xyz_flipped = xyz.transpose(1, 2).contiguous()
    inds = ref.furthest_point_sampling(xyz, npoint)
    new_xyz = ref.gather_points(xyz_flipped, inds).transpose(1, 2).contiguous()
    idx = ref.ball_query(new_xyz, xyz, radius, nsample)
    g = ref.group_points(xyz.transpose(1, 2).contiguous(), idx)
    g -= new_xyz.transpose(1, 2).unsqueeze(-1)
    g /= radius
    if feats is not None:
        g = torch.cat([g, ref.group_points(feats, idx)], dim=1)
    return inds, new_xyz, g


def our_front(xyz, feats, npoint, radius, nsample):
    inds, new_xyz = pointnet2_utils.fps_with_xyz(xyz, npoint)
    g, _, _ = pointnet2_utils.fused_query_and_group(xyz, new_xyz, feats, radius, nsample, True, True)
    return inds, new_xyz, g


pc = torch.from_numpy(scannet_like_batch(B, 80000, seed0=0)).to(dev)
xyz = pc[..., :3].contiguous()
feats = pc[..., 3:].transpose(1, 2).contiguous()
layers = [("SA1", 2048, 0.2, 64, 1), ("SA2", 1024, 0.4, 32, 128), ("SA3", 512, 0.8, 16, 256), ("SA4", 256, 1.2, 16, 256)]
rows = []
tot_r = tot_o = 0.0
cur_xyz, cur_f = xyz, feats
g = torch.Generator(device="cpu").manual_seed(0)
for name, npoint, radius, nsample, c_next in layers:
    tr, (ri, rx, rg) = timeit(lambda: ref_front(cur_xyz, cur_f, npoint, radius, nsample), iters=5 if name == "SA1" else 20)
    to, (oi, ox, og) = timeit(lambda: our_front(cur_xyz, cur_f, npoint, radius, nsample), iters=20)
    assert torch.equal(ri, oi) and torch.equal(rx, ox) and torch.equal(rg, og), name
    # per-op reference numbers
    t_fps_r, _ = timeit(lambda: ref.furthest_point_sampling(cur_xyz, npoint), iters=5 if name == "SA1" else 20)
    t_fps_o, _ = timeit(lambda: _ext.furthest_point_sampling(cur_xyz, npoint), iters=20)
    t_bq_r, _ = timeit(lambda: ref.ball_query(rx, cur_xyz, radius, nsample), iters=5 if name == "SA1" else 20)
    t_bq_o, _ = timeit(lambda: _ext.ball_query(rx, cur_xyz, radius, nsample), iters=20)
    rows.append((name, cur_xyz.shape[1], npoint, nsample, tr, to, t_fps_r, t_fps_o, t_bq_r, t_bq_o))
    tot_r += tr
    tot_o += to
    cur_xyz = rx
    cur_f = torch.randn(B, c_next, npoint, generator=g).to(dev)

print(f"# {B} scenes x 80000 points, B200, CUDA events; outputs verified bit-identical before timing")
print("| layer | N -> npoint (S) | reference front end ms | ours ms | x | ref FPS ms | our FPS ms | x | ref ball_query ms | our ball_query ms | x |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for name, n, m, s, tr, to, fr, fo, br, bo in rows:
    print(f"| {name} | {n} -> {m} ({s}) | {tr:.3f} | {to:.3f} | {tr / to:.1f} | {fr:.3f} | {fo:.3f} | {fr / fo:.1f} | {br:.3f} | {bo:.3f} | {br / bo:.1f} |")
print(f"| all four | | {tot_r:.3f} | {tot_o:.3f} | {tot_r / tot_o:.1f} | | | | | | |")
