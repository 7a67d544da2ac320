"""Data-parallel plumbing: one process per GPU, scenes sharded by batch index, no
data-path collective for inference, ONE bucketed gradient all-reduce per training step.

Replaces the reference's single-process nn.DataParallel (net_utils/utils.py:238: per-step parameter
broadcast + output gather + gradient reduce onto GPU 0).  BatchNorm statistics stay per rank, which matches
DataParallel's per-replica BN.  Backend: "nccl" on GPUs (NVLink 5 / NVSwitch), "gloo" in CPU tests.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment; returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(num_items, rank, world):
    """Contiguous shard [lo, hi) of `num_items` scenes (or objects) owned by `rank`; sizes differ by at most 1."""
    base, rem = divmod(num_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_gradients(params, world=None, average=True, bucket_dtype=None):
    """One flat all-reduce over every gradient (the single exchange step of the training path).
    Returns the number of bytes reduced."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    if bucket_dtype is not None:
        flat = flat.to(bucket_dtype)
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if average:
            flat = flat / world
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g).to(g.dtype))
        off += n
    return flat.numel() * flat.element_size()


def max_over_ranks(value, device=None):
    """max of a python float over all ranks (bench timing rule: the slowest rank defines the step time)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
