"""Run the ONet decoder alone (B objects x 32^3) a few times -- timing tool and target command for ncu captures.
usage: prof_decoder.py [objects=256] [iters=3] [precision=fp16] [cluster=2]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rfdnet_b200 import _lib, onet
from rfdnet_b200.synth import seeded_fill
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
prec = sys.argv[3] if len(sys.argv) > 3 else "fp16"
cluster = int(sys.argv[4]) if len(sys.argv) > 4 else 2
dev = torch.device("cuda:0")
_lib.check(_lib.load().rfd_onet_decode_set_cluster(cluster), "set_cluster")
dec = onet.DecoderCBatchNorm(dim=3, z_dim=32, c_dim=512).eval()
seeded_fill(dec, 31)
dec = dec.to(dev)
grid = onet.make_3d_grid(32, 1.1, dev)
c = torch.randn(B, 512, device=dev)
z = torch.zeros(B, 32, device=dev)
with torch.no_grad():
    for _ in range(iters):
        out = dec.decode(grid, z, c, precision=prec)
torch.cuda.synchronize()
_lib.TIMERS = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    e0.record()
    for _ in range(iters):
        out = dec.decode(grid, z, c, precision=prec)
    e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
kms = sum(s.elapsed_time(e) for n, s, e, w in _lib.TIMERS if n == "onet_decode") / iters
fl = B * 32768 * 1312768.0
print(f"decode B={B} {prec} cluster={cluster}: {ms:.3f} ms incl. cbn tables ({fl / ms / 1e9:.1f} TFLOP/s); "
      f"decode kernel alone {kms:.3f} ms = {fl / kms / 1e9:.1f} algorithmic TFLOP/s")
