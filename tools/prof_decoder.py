"""Run the ONet decoder alone (256 objects x 32^3) a few times -- target command for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rfdnet_b200 import onet
from rfdnet_b200.synth import seeded_fill
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
dec = onet.DecoderCBatchNorm(dim=3, z_dim=32, c_dim=512).eval()
seeded_fill(dec, 31)
dec = dec.to(dev)
grid = onet.make_3d_grid(32, 1.1, dev)
c = torch.randn(B, 512, device=dev)
z = torch.zeros(B, 32, device=dev)
with torch.no_grad():
    for _ in range(iters):
        out = dec.decode(grid, z, c)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    e0.record()
    out = dec.decode(grid, z, c)
    e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"decode B={B}: {ms:.3f} ms  -> {B * 32768 * 1312768.0 / ms / 1e9:.1f} TFLOP/s (incl. cbn tables)")
