"""BASELINE config 5 on hardware: the joint ISCNet training step (detection + SkipPropagation + ONet in train mode) on
the sm_100a point-cloud kernels, and the data-parallel exchange -- the bucketed, backward-overlapped gradient all-reduce
(rfdnet_b200/train.py) -- checked the way SURVEY.md 8e asks: gradients averaged over two ranks, each holding half of the
scenes, equal the single-process full-batch gradients."""
import os
import subprocess
import sys

import pytest
import torch

from rfdnet_b200 import train as T
from rfdnet_b200.synth import scannet_like_batch, seeded_fill

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = "cuda:0"


def test_joint_train_step_all_parameters_learn():
    torch.manual_seed(0)
    B, N, K = 2, 20000, 4
    model = T.JointTrainStep(boxes_per_scene=K)
    seeded_fill(model, 3)
    model = model.to(DEV).train()
    pc = torch.from_numpy(scannet_like_batch(B, N, seed0=40)).to(DEV)
    lab = T.synthetic_labels(B, N, K, 512, DEV, seed=1)
    tr = T.Trainer(model, lr=1e-4, bucket_bytes=4 << 20)
    assert len(tr.buckets.buckets) > 4
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    l0, parts = tr.step(pc, lab)
    for n, p in model.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
    dead = [n for n, p in model.named_parameters() if float(p.grad.abs().sum()) == 0.0]
    # every parameter takes part (the reference's cls-code path is disabled in ISCNet.yaml and has no parameters here)
    assert not dead, dead
    changed = sum(int(not torch.equal(before[n], p.detach())) for n, p in model.named_parameters())
    assert changed == len(before)
    l1, _ = tr.step(pc, lab)
    l2, _ = tr.step(pc, lab)
    assert all(torch.isfinite(x) for x in (l0, l1, l2))
    assert all(torch.isfinite(v) for v in parts.values())
    # gradients live in the flat buckets (no flatten / copy-back pass)
    b0 = tr.buckets.buckets[0]
    assert b0["params"][0].grad.data_ptr() == b0["buf"].data_ptr()


_WORKER = r"""
import os, sys, torch
sys.path.insert(0, %r)
import torch.distributions as dist
from rfdnet_b200 import dist as D, train as T
from rfdnet_b200.synth import scannet_like_batch, seeded_fill
backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"      # one GPU: both ranks share cuda:0, gloo moves the buckets
rank, world, local = D.init_from_env(backend)
dev = torch.device("cuda", local if backend == "nccl" else 0)
torch.cuda.set_device(dev)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
B, N, K, Tp = 4, 12000, 3, 256
model = T.JointTrainStep(boxes_per_scene=K)
seeded_fill(model, 9)
model = model.to(dev).eval()           # running-statistics BN: the loss is then a plain mean over scenes (shardable)
pc = torch.from_numpy(scannet_like_batch(B, N, seed0=70)).to(dev)
lab = T.synthetic_labels(B, N, K, Tp, dev, seed=2)
lab["vote_label_mask"] = torch.ones_like(lab["vote_label_mask"])       # equal denominators on every shard
eps = torch.randn(B * K, 32, generator=torch.Generator().manual_seed(5)).to(dev)
class FixedNoise(dist.Normal):
    def rsample(self, sample_shape=torch.Size()):
        return self.loc + self.scale * self.noise
def with_noise(model, e):
    orig = model.completion.infer_z
    def infer_z(*a, **k):
        q = orig(*a, **k)
        f = FixedNoise(q.loc, q.scale); f.noise = e
        return f
    model.completion.infer_z = infer_z
    return orig
lo, hi = D.shard_range(B, rank, world)
sl = lambda t: t[lo:hi].contiguous()
buckets = T.GradBuckets(list(model.parameters()), bucket_bytes=2 << 20)
buckets.zero()
orig = with_noise(model, eps[lo * K:hi * K])
loss, _ = model(sl(pc), {k: sl(v) for k, v in lab.items()})
loss.backward()
buckets.finish()                       # averaged over the two ranks
got = [p.grad.detach().clone() for p in model.parameters()]
model.completion.infer_z = orig
for p in model.parameters():
    p.grad = None
with_noise(model, eps)
loss_full, _ = model(pc, lab)
loss_full.backward()
worst = 0.0
for (n, p), g in zip(model.named_parameters(), got):
    ref = p.grad
    err = float((g - ref).abs().max()); scale = float(ref.abs().max())
    worst = max(worst, err / (scale + 1e-6))
    assert err <= 2e-3 * scale + 1e-5, (n, err, scale)
D.barrier()
sys.stdout.write("rank%%d-ok backend=%%s worst_rel=%%.2e\n" %% (rank, backend, worst)); sys.stdout.flush()
torch.distributed.destroy_process_group()
"""


@pytest.mark.timeout(600)
def test_two_rank_averaged_gradients_equal_full_batch(tmp_path):
    script = tmp_path / "worker_train.py"
    script.write_text(_WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541", OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                       capture_output=True, text=True, env=env, timeout=560)
    print(r.stdout[-1500:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("-ok") == 2, r.stdout
