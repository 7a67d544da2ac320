"""Training-path check (BASELINE config 5 shape: 8 ScanNet-like 80k-point scenes per GPU): the detection networks in
TRAIN mode (batch-statistics BatchNorm, autograd) run the reference's op sequence on the drop-in `_ext` kernels
(forward: FPS / ball query / group / 3-NN / interpolate; backward: the atomic scatter-add grads), followed by the
single flat gradient all-reduce (NCCL when launched with torchrun) and an Adam step.

The reference's DetectionLoss (models/loss.py) is outside the hot path; a surrogate scalar loss over
every head output is used so that every parameter and every backward kernel is exercised.  Gradients are checked
finite and non-zero.  Reports ms per step (max over ranks)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rfdnet_b200 import detection, dist as D
from rfdnet_b200.synth import scannet_like_batch, seeded_fill

rank, world, local = D.init_from_env("nccl")
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
net = detection.DetectionHotPath(1, 256)
seeded_fill(net, 0)
net = net.to(dev).train()
opt = torch.optim.Adam(net.parameters(), lr=1e-3)
pc = torch.from_numpy(scannet_like_batch(B, 80000, seed0=100 * rank)).to(dev)


def step():
    opt.zero_grad(set_to_none=True)
    ep, _ = net(pc)
    loss = (ep["objectness_scores"].pow(2).mean() + (ep["center"] - ep["seed_xyz"].mean(1, keepdim=True)).pow(2).mean()
            + ep["heading_scores"].pow(2).mean() + ep["heading_residuals_normalized"].pow(2).mean()
            + ep["size_scores"].pow(2).mean() + ep["size_residuals_normalized"].pow(2).mean()
            + ep["sem_cls_scores"].pow(2).mean() + 0.1 * ep["vote_xyz"].pow(2).mean())
    loss.backward()
    nbytes = D.allreduce_gradients(list(net.parameters()), world)
    opt.step()
    return float(loss.detach()), nbytes


loss0, nbytes = step()
grads = [p.grad for p in net.parameters() if p.grad is not None]
assert len(grads) == len(list(net.parameters())), "a parameter received no gradient"
assert all(torch.isfinite(g).all() for g in grads)
assert sum(float(g.abs().sum()) for g in grads) > 0
for _ in range(2):
    step()
torch.cuda.synchronize(); D.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 5
e0.record()
for _ in range(iters):
    l, _ = step()
e1.record()
torch.cuda.synchronize()
ms = D.max_over_ranks(e0.elapsed_time(e1) / iters, dev)
if rank == 0:
    print(f"train step (detection nets, surrogate loss): {B} scenes/GPU x {world} GPU(s): {ms:.1f} ms/step = "
          f"{B * world / ms * 1e3:.1f} scenes/s; all-reduce {nbytes / 1e6:.2f} MB; loss {loss0:.4f} -> {l:.4f}")
if world > 1:
    torch.distributed.destroy_process_group()
