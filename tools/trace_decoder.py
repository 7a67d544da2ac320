"""Timeline of the decoder's MMA <-> epilogue hand-offs (CTA 0, second tile), in SM cycles relative to layer start."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rfdnet_b200 import _lib, onet
from rfdnet_b200.synth import seeded_fill
dev = torch.device("cuda:0")
dec = onet.DecoderCBatchNorm(dim=3, z_dim=32, c_dim=512).eval(); seeded_fill(dec, 31); dec = dec.to(dev)
grid = onet.make_3d_grid(32, 1.1, dev)
c = torch.randn(256, 512, device=dev); z = torch.zeros(256, 32, device=dev)
with torch.no_grad():
    dec.decode(grid, z, c)
trace = torch.zeros(2 * 10 * 16, dtype=torch.int64, device=dev)
with torch.no_grad():
    dec.decode_traced(grid, z, c, trace)
torch.cuda.synchronize()
t = trace.cpu().view(2, 10, 16).numpy()
tile = t[1]
base = tile[0, 0]
print("second tile of CTA 0; cycles relative to the first MMA-ready of the tile")
print("layer | mma ready kp0..3 | mma issued kp0..3 | epi sees acc | epi published kp0..3")
for l in range(10):
    r = tile[l] - base
    print(l, "|", list(r[0:4]), "|", list(r[4:8]), "|", r[8], "|", list(r[9:13]))
print("tile period (layer-0 mma ready, tile 0 -> tile 1):", int(t[1, 0, 0] - t[0, 0, 0]), " sum of the 9 layer periods of tile 1:",
      int(tile[9, 0] - tile[0, 0]), " last layer's first mma ready -> next tile's first mma ready (tile 0 -> 1):", int(t[1, 0, 0] - t[0, 9, 0]))
print("layer period (mma ready kp0 -> next layer's):", [int(tile[l + 1, 0] - tile[l, 0]) for l in range(9)])
print("gap: last issue of layer l -> first ready of layer l+1:", [int(tile[l + 1, 0] - tile[l, 7]) for l in range(9)])
print("epilogue: acc seen -> panel0 published:", [int(tile[l, 9] - tile[l, 8]) for l in range(9)])
print("mma: last issue -> epilogue sees acc:", [int(tile[l, 8] - tile[l, 7]) for l in range(10)])
print("epilogue panel0 published -> mma ready (next layer):", [int(tile[l + 1, 0] - tile[l, 9]) for l in range(9)])
print("epilogue panel 0 breakdown: acc seen -> TMEM load done:", [int(tile[l, 13] - tile[l, 8]) for l in range(9)])
print("                              load done -> stores issued:", [int(tile[l, 14] - tile[l, 13]) for l in range(9)])
print("                              fence.proxy.async:", [int(tile[l, 15] - tile[l, 14]) for l in range(9)])
print("                              fence done -> published:", [int(tile[l, 9] - tile[l, 15]) for l in range(9)])
