"""CPU tests of the host side: C-ABI export table, module mirrors (state_dict compatibility, BN folding),
the model-level oracle against the committed golden fixtures, and the gloo world_size-2 sharding path."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from oracle import model_ref  # noqa: E402
from rfdnet_b200 import detection, mlp, onet, pointnet2_modules  # noqa: E402
from rfdnet_b200.synth import scannet_like_batch, seeded_fill, uniform_cloud  # noqa: E402


def test_cabi_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "rfdnet_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(rfd_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 25
    from rfdnet_b200 import _lib
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/rfdnet_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert lib.rfd_abi_version() == 4
    assert b"invalid" in lib.rfd_status_string(-1)
    # argument validation happens before any CUDA call, so it is testable without a GPU
    assert lib.rfd_furthest_point_sampling(None, 1, 0, 4, None, None) == -1
    assert lib.rfd_ball_query(None, None, 1, 10, 10, 0.1, 4, None, None) == -1
    assert lib.rfd_onet_decode(None, 0, 1, 128, None, None, 3, None, None, 0.0, None, None) == -1
    assert lib.rfd_onet_packed_bytes(1) == 10 * 4 * 256 * 128 == lib.rfd_onet_packed_bytes(2)
    assert lib.rfd_onet_packed_bytes(3) == 2 * 10 * 4 * 256 * 128
    assert lib.rfd_onet_aff_floats() == 2 * 11 * 2 * 256 + 256
    # chain-MLP plan (host-only): supported shapes report their packed size, unsupported ones 0
    assert lib.rfd_mlp_chain_packed_bytes(3, 512, 0, 256, 256, 0) > 0        # FP module, K0 streams in two rounds
    assert lib.rfd_mlp_chain_packed_bytes(3, 200, 0, 512, 0, 0) > 0          # one layer, two column blocks
    assert lib.rfd_mlp_chain_packed_bytes(3, 300, 0, 512, 0, 0) == 0         # second block would need non-resident A
    assert lib.rfd_mlp_chain_packed_bytes(3, 64, 0, 300, 64, 0) == 0         # hidden layer wider than 256


def test_no_oracle_import_in_product():
    """The product package must never import the oracle (or the reference)."""
    for root, _, files in os.walk(os.path.join(ROOT, "rfdnet_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "liboracle" not in src and "/root/reference" not in src.replace("/root/reference/", "REF/") or f.endswith((".cu", ".cuh", ".py"))


def test_ops_fail_loudly_without_gpu():
    from rfdnet_b200 import _ext
    with pytest.raises(RuntimeError, match="CPU not supported"):
        _ext.furthest_point_sampling(torch.zeros(1, 8, 3), 2)
    with pytest.raises(RuntimeError, match="contiguous"):
        _ext.ball_query(torch.zeros(1, 4, 3).transpose(0, 1).transpose(0, 1)[:, :, :].permute(0, 2, 1), torch.zeros(1, 8, 3), 0.1, 2)
    with pytest.raises(RuntimeError, match="int tensor"):
        _ext.gather_points(torch.zeros(1, 3, 8), torch.zeros(1, 2, dtype=torch.int64))


def test_state_dict_keys_match_reference(golden):
    from rfdnet_b200 import stn_group
    mods = {"backbone": detection.Pointnet2Backbone(1), "voting": detection.VotingModule(1, 256),
            "detection": detection.ProposalModule(), "decoder": onet.DecoderCBatchNorm(z_dim=32, c_dim=512),
            "stn_group": stn_group.STN_Group(radius=1.0, nsample=1024, use_xyz=False)}
    for name, m in mods.items():
        mine = [f"{k}:{tuple(v.shape)}" for k, v in m.state_dict().items()]
        ref = [str(x) for x in golden[f"keys_{name}"]]
        assert sorted(mine) == sorted(ref), name


def test_fold_conv_bn_matches_torch_eval():
    torch.manual_seed(0)
    conv, bn = torch.nn.Conv2d(7, 5, 1, bias=False), torch.nn.BatchNorm2d(5)
    seeded_fill(bn, 3)
    bn.eval()
    x = torch.randn(2, 7, 11, 3)
    W, s, t = mlp.fold_conv_bn(conv, bn)
    ref = bn(conv(x))
    got = torch.einsum("oc,bcnm->bonm", W, x) * s.view(1, -1, 1, 1) + t.view(1, -1, 1, 1)
    assert torch.allclose(got, ref.detach(), atol=1e-5)
    conv1 = torch.nn.Conv1d(7, 5, 1)
    W, s, t = mlp.fold_conv_bn(conv1, None)
    assert torch.allclose(torch.einsum("oc,bcn->bon", W, x[..., 0]) * s.view(1, -1, 1) + t.view(1, -1, 1),
                          conv1(x[..., 0]).detach(), atol=1e-5)
    layers = mlp.fold_sequential(pointnet2_modules.build_shared_mlp([4, 8, 16]))
    assert [tuple(l[0].shape) for l in layers] == [(8, 4), (16, 8)] and all(l[3] for l in layers)


def test_model_ref_golden_sa_fp(golden):
    """oracle/model_ref against fixtures produced by the reference's own modules (CPU => true division)."""
    sa = pointnet2_modules.PointnetSAModuleVotes(npoint=128, radius=0.3, nsample=16, mlp=[5, 32, 32, 64],
                                                 use_xyz=True, normalize_xyz=True).eval()
    seeded_fill(sa, 11)
    g = torch.Generator().manual_seed(3)
    xyz = torch.from_numpy(uniform_cloud(2, 1024, seed=5))
    feats = torch.randn(2, 5, 1024, generator=g)
    sd = {"sa." + k: v for k, v in sa.state_dict().items()}
    nx, nf, ind = model_ref.sa_module(xyz, feats, sd, "sa", 128, 0.3, 16, recip=False)
    assert np.array_equal(ind.numpy(), golden["sa_inds"])
    assert np.array_equal(nx.numpy(), golden["sa_new_xyz"])
    assert np.allclose(nf.numpy(), golden["sa_new_features"], atol=1e-6, rtol=1e-6)
    fp = pointnet2_modules.PointnetFPModule(mlp=[64 + 16, 64, 32]).eval()
    seeded_fill(fp, 12)
    unk = torch.from_numpy(uniform_cloud(2, 300, seed=6))
    kn = torch.from_numpy(uniform_cloud(2, 64, seed=7))
    uf, kf = torch.randn(2, 16, 300, generator=g), torch.randn(2, 64, 64, generator=g)
    sd = {"fp." + k: v for k, v in fp.state_dict().items()}
    assert np.allclose(model_ref.fp_module(unk, kn, uf, kf, sd, "fp").numpy(), golden["fp_out"], atol=1e-6, rtol=1e-6)


def test_model_ref_golden_decoder_and_grid(golden):
    dec = onet.DecoderCBatchNorm(dim=3, z_dim=32, c_dim=512).eval()
    seeded_fill(dec, 31)
    grid = model_ref.make_3d_grid(32, 1.1)
    assert np.array_equal(grid[:32, 2].numpy(), golden["grid32_axis"])
    p = grid[torch.from_numpy(golden["dec_sel"]).long()].unsqueeze(0).expand(3, -1, -1).contiguous()
    c = torch.from_numpy(golden["dec_c"])
    sd = dict(dec.state_dict())
    out = model_ref.decoder(p, torch.zeros(3, 32), c, sd)
    assert np.allclose(out.numpy(), golden["dec_logits"], atol=1e-5, rtol=1e-5)
    out = model_ref.decoder(p, torch.from_numpy(golden["dec_z2"]), c, sd)
    assert np.allclose(out.numpy(), golden["dec_logits_z"], atol=1e-5, rtol=1e-5)
    # the mirror module's own PyTorch (training-path) forward is the same function
    with torch.no_grad():
        assert torch.allclose(dec(p, torch.zeros(3, 32), c), torch.from_numpy(golden["dec_logits"]), atol=1e-5, rtol=1e-5)


@pytest.mark.timeout(600)
def test_model_ref_golden_detection(golden):
    net = detection.DetectionHotPath(1, 256).eval()
    seeded_fill(net.backbone, 21)
    seeded_fill(net.voting, 22)
    seeded_fill(net.detection, 23)
    sd = dict(net.state_dict())
    pc = torch.from_numpy(scannet_like_batch(1, 20000, seed0=100))
    ep = model_ref.backbone(pc, sd, recip=False)
    assert np.array_equal(ep["sa1_inds"].numpy(), golden["det_sa1_inds"])
    assert bool(golden["det_sa2_inds_is_arange"]) == bool(torch.equal(ep["sa2_inds"][0].long(), torch.arange(1024)))
    assert np.allclose(ep["sa4_features"].numpy()[:, :, :32], golden["det_sa4_features"], atol=2e-5, rtol=1e-5)
    assert np.allclose(ep["fp2_features"].numpy()[:, :, :64], golden["det_fp2_features"], atol=2e-5, rtol=1e-5)
    vx, vf = model_ref.voting(ep["fp2_xyz"], ep["fp2_features"], sd)
    assert np.allclose(vx.numpy()[:, :128], golden["det_vote_xyz"], atol=2e-5, rtol=1e-5)
    ax, inds, netp = model_ref.proposal(vx, vf, sd, recip=False)
    assert np.array_equal(inds.numpy(), golden["det_agg_inds"])
    sc = netp.transpose(2, 1)
    assert np.allclose(sc[:, :, 0:2].numpy(), golden["det_objectness"], atol=1e-4, rtol=1e-4)
    assert np.allclose((ax + sc[:, :, 2:5]).numpy(), golden["det_center"], atol=1e-4, rtol=1e-4)
    assert np.allclose(sc[:, :, -8:].numpy(), golden["det_sem_cls"], atol=1e-4, rtol=1e-4)


def test_sample_uniformly_whole_tensor_form():
    """QueryAndGroup(sample_uniformly=True), pointnet2_utils.py:321-330: distinct indices kept (ascending), the rest of
    every row drawn from them."""
    from rfdnet_b200.pointnet2_utils import _resample_uniformly
    g = torch.Generator().manual_seed(0)
    idx = torch.randint(0, 6, (3, 7, 16), generator=g, dtype=torch.int32)
    idx[0, 0] = 4
    idx[1, 2] = torch.arange(16, dtype=torch.int32)
    out, cnt = _resample_uniformly(idx)
    assert out.shape == idx.shape and out.dtype == torch.int32 and cnt.shape == (3, 7)
    for b in range(3):
        for m in range(7):
            u = torch.unique(idx[b, m])
            n = len(u)
            assert cnt[b, m] == n and torch.equal(out[b, m, :n].long(), u.long())
            assert set(out[b, m, n:].tolist()) <= set(u.tolist())


def test_shard_range():
    from rfdnet_b200.dist import shard_range
    for n in (0, 1, 7, 8, 64, 257):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


_GLOO_WORKER = r"""
import os, sys, torch
sys.path.insert(0, %r)
from rfdnet_b200 import dist as D
rank, world, _ = D.init_from_env("gloo")
torch.manual_seed(0)
net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 1))
data = torch.randn(8, 6); tgt = torch.randn(8, 1)
lo, hi = D.shard_range(8, rank, world)
loss = ((net(data[lo:hi]) - tgt[lo:hi]) ** 2).sum() / 8 * world   # so that the rank-average equals the full-batch mean loss grad
loss.backward()
nbytes = D.allreduce_gradients(list(net.parameters()), world)
net2 = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 1))
net2.load_state_dict(net.state_dict())
(((net2(data) - tgt) ** 2).sum() / 8).backward()
for a, b in zip(net.parameters(), net2.parameters()):
    assert torch.allclose(a.grad, b.grad, atol=1e-6), (a.grad, b.grad)
assert nbytes == sum(p.numel() for p in net.parameters()) * 4
t = D.max_over_ranks(float(rank + 1))
assert t == float(world)
D.barrier()
sys.stdout.write("rank" + str(rank) + "-ok\n"); sys.stdout.flush()
"""


_GLOO_BUCKETS_WORKER = r"""
import os, sys, torch
sys.path.insert(0, %r)
from rfdnet_b200 import dist as D
from rfdnet_b200.train import GradBuckets
rank, world, _ = D.init_from_env("gloo")
torch.manual_seed(0)
def make():
    return torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 9), torch.nn.ReLU(),
                               torch.nn.Linear(9, 1), torch.nn.Linear(1, 1))
net = make()
unused = torch.nn.Parameter(torch.ones(3))            # a parameter that never receives a gradient
data = torch.randn(8, 6); tgt = torch.randn(8, 1)
buckets = GradBuckets(list(net.parameters()) + [unused], bucket_bytes=64)    # tiny buckets: several per step
assert len(buckets.buckets) >= 3
net2 = make(); net2.load_state_dict(net.state_dict())
lo, hi = D.shard_range(8, rank, world)
for step in range(2):                                   # twice: counters / views must re-arm
    buckets.zero()
    loss = ((net(data[lo:hi]) - tgt[lo:hi]) ** 2).sum() / 8 * world
    loss.backward()                                     # hooks launch the per-bucket all-reduces during backward
    buckets.finish()
    net2.zero_grad(set_to_none=True)
    (((net2(data) - tgt) ** 2).sum() / 8).backward()
    for a, b in zip(net.parameters(), net2.parameters()):
        assert torch.allclose(a.grad, b.grad, atol=1e-6), (a.grad, b.grad)
    assert float(unused.grad.abs().sum()) == 0.0
assert buckets.nbytes == (sum(p.numel() for p in net.parameters()) + 3) * 4
# gradients are views into the flat buckets (no flatten / copy-back)
b0 = buckets.buckets[0]
assert b0["params"][0].grad.data_ptr() == b0["buf"].data_ptr()
D.barrier()
sys.stdout.write("rank" + str(rank) + "-ok\n"); sys.stdout.flush()
"""


@pytest.mark.timeout(300)
def test_gloo_world2_bucketed_overlapped_allreduce(tmp_path):
    """GradBuckets (rfdnet_b200/train.py): the averaged 2-rank gradients equal the single-process full-batch gradients."""
    script = tmp_path / "worker_b.py"
    script.write_text(_GLOO_BUCKETS_WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, env=env, timeout=280)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("-ok") == 2, r.stdout


@pytest.mark.timeout(300)
def test_gloo_world2_gradient_allreduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)],
                       capture_output=True, text=True, env=env, timeout=280)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("-ok") == 2 and "rank0" in r.stdout and "rank1" in r.stdout, r.stdout


def test_c_abi_from_plain_c(lib, tmp_path):
    """Compile a C99 program against include/rfdnet_b200.h with gcc, link it to librfdnet_b200.so and run it."""
    from rfdnet_b200 import _lib
    exe = tmp_path / "abi_check"
    libdir = os.path.dirname(_lib.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
           "-isystem", "/usr/local/cuda/include", os.path.join(ROOT, "tests", "c", "abi_check.c"), "-o", str(exe),
           "-L", libdir, "-lrfdnet_b200", "-L", "/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + libdir,
           "-Wl,-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "abi_check ok" in r.stdout, r.stdout + r.stderr


def test_dropin_install_registers_ext():
    """after dropin.install(), `import pointnet2_ops._ext` resolves to rfdnet_b200._ext and
    exposes the reference's nine function names (bindings.cpp:6-19)."""
    import importlib
    saved = {k: sys.modules.get(k) for k in ("pointnet2_ops", "pointnet2_ops._ext")}
    try:
        sys.modules.pop("pointnet2_ops", None)
        sys.modules.pop("pointnet2_ops._ext", None)
        from rfdnet_b200 import dropin
        ext = dropin.install()
        mod = importlib.import_module("pointnet2_ops._ext")
        assert mod is ext
        for name in ("gather_points", "gather_points_grad", "furthest_point_sampling", "three_nn", "three_interpolate",
                     "three_interpolate_grad", "ball_query", "group_points", "group_points_grad"):
            assert callable(getattr(mod, name)), name
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
