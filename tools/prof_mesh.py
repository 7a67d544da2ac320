"""Time rfd_extract_mesh on decoder logits (256 objects x 32^3): python tools/prof_mesh.py [objects=256]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rfdnet_b200 import generator, onet
from rfdnet_b200.synth import seeded_fill
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
dec = onet.DecoderCBatchNorm(dim=3, z_dim=32, c_dim=512).eval(); seeded_fill(dec, 31); dec = dec.to(dev)
grid = onet.make_3d_grid(32, 1.1, dev)
c = torch.randn(B, 512, device=dev, generator=torch.Generator(device=dev).manual_seed(7)); z = torch.zeros(B, 32, device=dev)
with torch.no_grad():
    lg = dec.decode(grid, z, c)
pools = (torch.empty((B * 24576, 3), device=dev), torch.empty((B * 49152, 3), dtype=torch.int32, device=dev))
for _ in range(2):
    mb = generator.extract_meshes(lg, 32, pools=pools)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    mb = generator.extract_meshes(lg, 32, pools=pools)
e1.record()
torch.cuda.synchronize()
v, t, r = mb.to_host()
print(f"extract_mesh B={B}: {e0.elapsed_time(e1) / 5 * 1e3:.1f} us per call; {len(v)} vertices, {len(t)} triangles "
      f"({(v.nbytes + t.nbytes) / 1e6:.1f} MB vs logits {lg.numel() * 4 / 1e6:.1f} MB)")
