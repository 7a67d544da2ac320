"""SkipPropagation.generate for the 256 proposals of one 80k-point scene -- timing tool and target command for ncu launch
lists:  python tools/prof_skip.py [precision=x3] [iters=3]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rfdnet_b200 import completion
from rfdnet_b200.synth import scannet_like_batch, seeded_fill
prec = sys.argv[1] if len(sys.argv) > 1 else "x3"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
sp = completion.SkipPropagation(input_feature_dim=1, c_dim=512, hidden_dim=512).eval()
seeded_fill(sp, 17)
sp = sp.to(dev)
sp.fast_precision = None if prec == "torch" else prec
pc = torch.from_numpy(scannet_like_batch(1, 80000, seed0=3)).to(dev)
g = torch.Generator().manual_seed(0)
sel = torch.randint(0, 80000, (256,), generator=g)
box_xyz = (pc[:, sel, :3] + 0.05).contiguous()
heading = (torch.rand(1, 256, generator=g) * 6.28).to(dev)
box_feat = torch.randn(1, 128, 256, generator=g).to(dev)
with torch.no_grad():
    for _ in range(2):
        sp.generate(box_xyz, heading, box_feat, pc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = sp.generate(box_xyz, heading, box_feat, pc)
    e1.record()
    torch.cuda.synchronize()
print(f"SkipPropagation.generate 256 proposals, {prec}: {e0.elapsed_time(e1) / iters:.2f} ms")
