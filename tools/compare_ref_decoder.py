"""ONet decoder: this library's tcgen05 kernel vs the reference's PyTorch op sequence (cuDNN/cuBLAS) on the same
GPU, driven like Generator3D.eval_points (generator.py:123-143: one object of 32^3 points per call, `.cpu()` after each)
and batched.  Measurement infrastructure."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rfdnet_b200 import onet
from rfdnet_b200.synth import seeded_fill

dev = torch.device("cuda:0")
nobj = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dec = onet.DecoderCBatchNorm(dim=3, z_dim=32, c_dim=512).eval()
seeded_fill(dec, 31)
dec = dec.to(dev)
grid = onet.make_3d_grid(32, 1.1, dev)
c = torch.randn(nobj, 512, device=dev)
z = torch.zeros(nobj, 32, device=dev)


def wall(fn, iters):
    fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / iters


def ref_loop():  # generator.py:71-74,131-141
    out = []
    with torch.no_grad():
        for o in range(nobj):
            out.append(dec.forward_reference(grid.unsqueeze(0), z[o:o + 1], c[o:o + 1]).squeeze(0).cpu())
    return out


def ref_batched(chunk=16):
    with torch.no_grad():
        return [dec.forward_reference(grid.unsqueeze(0).expand(min(chunk, nobj - o), -1, -1), z[o:o + chunk], c[o:o + chunk])
                for o in range(0, nobj, chunk)]


def ours():
    with torch.no_grad():
        return dec.decode(grid, z, c).cpu()


print(f"# {nobj} objects x 32^3 query points, B200, wall clock incl. the device->host copy of the logits")
for tf32 in (True, False):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    t = wall(ref_loop, 1)
    print(f"reference op sequence, per-object loop + .cpu(), tf32={tf32}: {t * 1e3:.1f} ms ({t / nobj * 1e3:.2f} ms/object)")
    t = wall(ref_batched, 1)
    print(f"reference op sequence, 16 objects per call, on device,  tf32={tf32}: {t * 1e3:.1f} ms")
t = wall(ours, 3)
print(f"rfdnet_b200 decode (one call, default tcgen05 mode) + .cpu(): {t * 1e3:.1f} ms")
a = torch.cat([x.view(1, -1) for x in ref_loop()[:8]]).to(dev)
b = dec.decode(grid, z[:8], c[:8])
print(f"max |logit difference| vs the fp32 reference sequence on 8 objects: {float((a - b).abs().max()):.3e} (scale {float(a.abs().max()):.2f})")
