"""Time rfd_query_and_group on the five SA shapes of the benchmark step (B scenes of 80k points): algorithmic bytes
(12N + 12M + 4CN + 4(3+C)MS per scene) / CUDA-event time, back-to-back launches after warm-up.
  python tools/prof_qg.py [B] [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json

import torch

from bench import graph_time
from rfdnet_b200 import pointnet2_utils as pu
from rfdnet_b200.synth import scannet_like_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
ITERS = int(sys.argv[2]) if len(sys.argv) > 2 else 30
dev = torch.device("cuda:0")
pc = torch.from_numpy(scannet_like_batch(B, 80000, seed0=0)).to(dev)
xyz = pc[..., :3].contiguous()
LAYERS = [("SA1", 2048, 0.2, 64, 1), ("SA2", 1024, 0.4, 32, 128), ("SA3", 512, 0.8, 16, 256), ("SA4", 256, 1.2, 16, 256)]
peak = 6536.7
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p)).get("hbm_gbs", peak)
cur = xyz
tot_b = tot_t = 0.0
rows = []
shapes = []
for name, M, r, S, C in LAYERS:
    _, new_xyz = pu.fps_with_xyz(cur, M)
    shapes.append((name, cur, new_xyz, r, S, C))
    cur = new_xyz
# vote aggregation: 1024 votes -> 256 clusters, r = 0.3, S = 16, C = 256 (votes ~ seeds = SA2 points)
votes = shapes[1][2]
_, vq = pu.fps_with_xyz(votes, 256)
shapes.append(("vote-agg", votes, vq, 0.3, 16, 256))
for name, src, q, r, S, C in shapes:
    N, M = src.shape[1], q.shape[1]
    g = torch.Generator(device=dev).manual_seed(1)
    feats = torch.randn(B, C, N, device=dev, generator=g)
    if ITERS > 1:
        ms = graph_time(lambda: pu.fused_query_and_group(src, q, feats, r, S, True, True), ITERS)
    else:
        for _ in range(4):
            pu.fused_query_and_group(src, q, feats, r, S, True, True)
        torch.cuda.synchronize()
        ms = 1.0
    nbytes = B * (12 * N + 12 * M + 4 * C * N + 4 * (3 + C) * M * S)
    gbs = nbytes / (ms * 1e-3) / 1e9
    tot_b += nbytes
    tot_t += ms
    print(f"{name:9s} N={N:6d} M={M:5d} S={S:3d} C={C:4d}: {ms * 1e3:8.1f} us  {nbytes / 1e6:7.2f} MB  {gbs:8.1f} GB/s  "
          f"= {gbs / peak:.3f} of HBM peak ({peak:.0f})")
print(f"aggregate: {tot_b / 1e6:.1f} MB in {tot_t * 1e3:.1f} us = {tot_b / (tot_t * 1e-3) / 1e9:.1f} GB/s = "
      f"{tot_b / (tot_t * 1e-3) / 1e9 / peak:.3f} of HBM peak")
