"""torchrun --nproc-per-node N tools/allreduce_check.py -- the single exchange step of the training path
(SURVEY.md section 8e): one flat NCCL all-reduce over the gradients of the whole ISCNet-sized parameter set.
Checks the averaged gradients against the analytic value and reports the bus bandwidth."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from rfdnet_b200 import dist as D

rank, world, local = D.init_from_env("nccl")
dev = torch.device("cuda", local)
for nparam, name in ((950_000, "detection phase (0.95 M params)"), (15_200_000, "completion phase (15.2 M params)")):
    p = torch.nn.Parameter(torch.zeros(nparam, device=dev))
    p.grad = torch.full((nparam,), float(rank + 1), device=dev)
    nbytes = D.allreduce_gradients([p], world)
    expect = sum(range(1, world + 1)) / world
    assert torch.allclose(p.grad, torch.full_like(p.grad, expect)), (float(p.grad[0]), expect)
    for _ in range(5):
        D.allreduce_gradients([p], world)
    torch.cuda.synchronize(); D.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    iters = 20
    for _ in range(iters):
        D.allreduce_gradients([p], world)
    e1.record()
    torch.cuda.synchronize()
    ms = D.max_over_ranks(e0.elapsed_time(e1) / iters, dev)
    busbw = nbytes * 2 * (world - 1) / world / (ms * 1e-3) / 1e9 if world > 1 else 0.0
    if rank == 0:
        print(f"all-reduce {name}: {nbytes / 1e6:.1f} MB, {ms * 1e3:.1f} us per step (incl. flatten/unflatten), "
              f"bus bandwidth {busbw:.1f} GB/s, world {world}")
if world > 1:
    dist.destroy_process_group()
