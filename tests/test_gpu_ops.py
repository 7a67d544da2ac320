"""GPU parity tests of the nine `_ext` operators + the fused kernels, through the C ABI, against the CPU oracle
(bit-exact for indices / copies, tolerance for atomics) and -- when oracle/_ref/_ref_ext.so travelled to the box --
against the UNMODIFIED reference CUDA kernels recompiled for sm_100."""
import numpy as np
import pytest
import torch

import oracle
from oracle import build_ref, model_ref
from rfdnet_b200 import _ext, pointnet2_utils
from rfdnet_b200.synth import scannet_like_batch, tricky_cloud, uniform_cloud

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def gather_xyz(cloud, idx):
    return np.take_along_axis(cloud, idx[..., None].astype(np.int64).repeat(3, -1), 1)


@pytest.fixture(scope="module")
def ref_ext():
    m = build_ref.load()
    if m is None:
        pytest.skip("oracle/_ref/_ref_ext.so not present")
    return m


@pytest.mark.parametrize("cloud,m", [
    (uniform_cloud(2, 4096, seed=0), 512),      # config 1
    (tricky_cloud(4096, seed=1), 512),          # duplicates + points inside the skip radius
    (uniform_cloud(3, 37, seed=2), 20),         # N < 512: reference block size 32
    (uniform_cloud(1, 511, seed=3), 64),
    (uniform_cloud(2, 1000, seed=4), 1000),     # m == N
    (uniform_cloud(1, 8, seed=5), 12),          # m > N
    (uniform_cloud(1, 1, seed=6), 3),
    (np.zeros((1, 64, 3), np.float32), 5),      # every point skipped -> zeros
    (uniform_cloud(2, 9000, seed=7), 300),      # cluster path, small
    (np.random.default_rng(8).integers(-3, 4, (2, 20000, 3)).astype(np.float32), 200),  # lattice: massive ties
])
def test_fps_bit_exact_vs_oracle(cloud, m):
    got = _ext.furthest_point_sampling(cu(cloud), m).cpu().numpy()
    assert np.array_equal(got, oracle.furthest_point_sampling(cloud, m))


def test_fps_golden(golden):
    for tag, cloud in (("uniform", uniform_cloud(2, 4096, seed=0)), ("tricky", tricky_cloud(4096, seed=1))):
        got = _ext.furthest_point_sampling(cu(cloud), 512).cpu().numpy()
        assert np.array_equal(got, golden[f"c1_{tag}_fps"])


def test_fps_80k_scene_vs_oracle_and_properties():
    pc = scannet_like_batch(2, 80000, seed0=0)[..., :3].copy()
    got = _ext.furthest_point_sampling(cu(pc), 2048).cpu().numpy()
    assert np.array_equal(got[:1], oracle.furthest_point_sampling(pc[:1], 2048))
    for b in range(2):
        assert got[b, 0] == 0 and len(set(got[b].tolist())) == 2048
    # FPS of an FPS-ordered prefix is the identity (pointnet2backbone.py:104 comment) -- size-independent property
    sub = gather_xyz(pc, got)
    again = _ext.furthest_point_sampling(cu(sub), 1024).cpu().numpy()
    assert np.array_equal(again, np.tile(np.arange(1024, dtype=np.int32), (2, 1)))


@pytest.mark.parametrize("N,M,r,S", [(4096, 512, 0.2, 32), (4096, 512, 0.4, 32), (300, 77, 0.5, 16), (50, 9, 0.05, 8),
                                     (2048, 1024, 0.4, 32), (5000, 10, 1.0, 1024)])
def test_ball_query_bit_exact(N, M, r, S):
    cloud = uniform_cloud(2, N, seed=N)
    q = cloud[:, :M].copy()
    q[:, -1] = 50.0  # a query with an empty ball
    got = _ext.ball_query(cu(q), cu(cloud), r, S).cpu().numpy()
    assert np.array_equal(got, oracle.ball_query(q, cloud, r, S))


def test_ball_query_golden(golden):
    for tag, cloud in (("uniform", uniform_cloud(2, 4096, seed=0)), ("tricky", tricky_cloud(4096, seed=1))):
        q = gather_xyz(cloud, golden[f"c1_{tag}_fps"])
        assert np.array_equal(_ext.ball_query(cu(q), cu(cloud), 0.2, 32).cpu().numpy(), golden[f"c1_{tag}_bq02"])
        assert np.array_equal(_ext.ball_query(cu(q), cu(cloud), 0.4, 32).cpu().numpy(), golden[f"c1_{tag}_bq04"])


def test_ball_query_80k_properties():
    pc = scannet_like_batch(1, 80000, seed0=3)[..., :3].copy()
    fps = _ext.furthest_point_sampling(cu(pc), 2048)
    q = torch.gather(cu(pc), 1, fps.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    idx = _ext.ball_query(q, cu(pc), 0.2, 64).cpu().numpy()[0]
    qn = q.cpu().numpy()[0]
    d = np.linalg.norm(pc[0][idx] - qn[:, None, :], axis=-1)
    assert (d < 0.2 + 1e-6).all()                      # every returned neighbour is inside the ball
    for row in idx[:256]:
        u = row[: len(np.unique(row))] if len(np.unique(row)) < 64 else row
        assert (np.diff(u) > 0).all()                  # ascending index order before the padding
        assert (row[len(u):] == row[0]).all()          # padding repeats the first hit
    assert np.array_equal(idx[:64], oracle.ball_query(qn[None, :64], pc, 0.2, 64)[0])


@pytest.mark.parametrize("C,N,M,S", [(5, 100, 17, 4), (64, 4096, 512, 32), (3, 2048, 1024, 1)])
def test_group_and_gather(C, N, M, S):
    rng = np.random.default_rng(C)
    pts = rng.normal(size=(2, C, N)).astype(np.float32)
    idx = rng.integers(0, N, (2, M, S)).astype(np.int32)
    assert np.array_equal(_ext.group_points(cu(pts), cu(idx)).cpu().numpy(), oracle.group_points(pts, idx))
    assert np.array_equal(_ext.gather_points(cu(pts), cu(idx[:, :, 0])).cpu().numpy(),
                          oracle.gather_points(pts, idx[:, :, 0]))
    go = rng.normal(size=(2, C, M, S)).astype(np.float32)
    got = _ext.group_points_grad(cu(go), cu(idx), N).cpu().numpy()
    assert np.allclose(got, oracle.group_points_grad(go, idx, N), atol=1e-4, rtol=1e-4)  # atomics: order differs
    got = _ext.gather_points_grad(cu(go[..., 0].copy()), cu(idx[:, :, 0].copy()), N).cpu().numpy()
    assert np.allclose(got, oracle.gather_points_grad(go[..., 0], idx[:, :, 0], N), atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("n,m", [(512, 256), (1024, 512), (700, 3), (10, 2), (33, 1500)])
def test_three_nn_and_interpolate(n, m):
    unk, kn = uniform_cloud(2, n, seed=n), uniform_cloud(2, m, seed=m + 1)
    kn[:, 1] = kn[:, 0]  # exact duplicate -> tie
    d2, ix = _ext.three_nn(cu(unk), cu(kn))
    rd2, rix = oracle.three_nn(unk, kn)
    assert np.array_equal(ix.cpu().numpy(), rix)
    assert np.array_equal(d2.cpu().numpy(), rd2)
    rng = np.random.default_rng(n)
    feats = rng.normal(size=(2, 16, m)).astype(np.float32)
    w = rng.uniform(0.1, 1, (2, n, 3)).astype(np.float32)
    out = _ext.three_interpolate(cu(feats), cu(rix), cu(w)).cpu().numpy()
    assert np.array_equal(out, oracle.three_interpolate(feats, rix, w))  # same FMA order -> bit exact
    go = rng.normal(size=(2, 16, n)).astype(np.float32)
    gg = _ext.three_interpolate_grad(cu(go), cu(rix), cu(w), m).cpu().numpy()
    assert np.allclose(gg, oracle.three_interpolate_grad(go, rix, w, m), atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("N,M,C,r,S,norm", [(4096, 512, 0, 0.2, 32, True), (4096, 512, 64, 0.2, 32, True),
                                            (2048, 1024, 128, 0.4, 32, True), (1024, 300, 7, 0.3, 16, False),
                                            (512, 256, 256, 1.2, 16, True), (333, 13, 5, 0.25, 64, True),
                                            (6000, 200, 20, 0.2, 48, True),      # two staged chunks, S not a power of 2
                                            (5000, 33, 130, 0.5, 1024, False),   # STN_Group-sized nsample, C % 4 != 0
                                            (2049, 100, 9, 0.3, 8, True),        # scene 1 not 16-byte aligned: no bulk copy
                                            (1024, 1024, 256, 0.3, 16, True), (700, 5, 33, 0.4, 3, True)])
def test_fused_query_and_group_bit_exact(N, M, C, r, S, norm):
    cloud = uniform_cloud(2, N, seed=N + C)
    q = cloud[:, :M].copy()
    q[:, -1] = 40.0
    feats = None if C == 0 else np.random.default_rng(C).normal(size=(2, C, N)).astype(np.float32)
    out, gxyz, idx = pointnet2_utils.fused_query_and_group(cu(cloud), cu(q), None if feats is None else cu(feats), r, S,
                                                           True, norm, ret_grouped_xyz=True, ret_idx=True)
    rf, rg, ri = model_ref.query_and_group(torch.from_numpy(cloud), torch.from_numpy(q),
                                           None if feats is None else torch.from_numpy(feats), r, S, True, norm,
                                           recip=True)
    assert np.array_equal(idx.cpu().numpy(), ri.numpy())
    assert torch.equal(gxyz.cpu(), rg)
    assert torch.equal(out.cpu(), rf)


def test_autograd_functions_match_reference_semantics():
    """gradients through the drop-in Functions (pointnet2_utils.py:68-101,139-240) equal dense torch autograd."""
    torch.manual_seed(0)
    f = torch.randn(2, 6, 50, device=DEV, requires_grad=True)
    idx = torch.randint(0, 50, (2, 9, 4), device=DEV, dtype=torch.int32)
    out = pointnet2_utils.grouping_operation(f, idx)
    ref = torch.gather(f.unsqueeze(2).expand(-1, -1, 9, -1), 3, idx.long().unsqueeze(1).expand(-1, 6, -1, -1))
    assert torch.equal(out, ref)
    g = torch.randn_like(out)
    (ga,) = torch.autograd.grad(out, f, g)
    (gb,) = torch.autograd.grad(ref, f, g)
    assert torch.allclose(ga, gb, atol=1e-5)
    w = torch.rand(2, 11, 3, device=DEV)
    i3 = torch.randint(0, 50, (2, 11, 3), device=DEV, dtype=torch.int32)
    o2 = pointnet2_utils.three_interpolate(f, i3, w)
    r2 = sum(torch.gather(f, 2, i3[:, :, t].long().unsqueeze(1).expand(-1, 6, -1)) * w[:, :, t].unsqueeze(1) for t in range(3))
    assert torch.allclose(o2, r2, atol=1e-6)
    g2 = torch.randn_like(o2)
    assert torch.allclose(torch.autograd.grad(o2, f, g2)[0], torch.autograd.grad(r2, f, g2)[0], atol=1e-5)


# ------------------------------------------------------------------ vs the unmodified reference CUDA kernels
def test_vs_reference_cuda_kernels(ref_ext):
    for cloud, m in ((uniform_cloud(2, 4096, seed=0), 512), (tricky_cloud(4096, seed=1), 512),
                     (scannet_like_batch(1, 80000, seed0=1)[..., :3].copy(), 2048), (uniform_cloud(2, 300, seed=9), 64)):
        x = cu(cloud)
        a = _ext.furthest_point_sampling(x, m)
        b = ref_ext.furthest_point_sampling(x, m)
        assert torch.equal(a, b), "FPS differs from the reference CUDA kernel"
        q = torch.gather(x, 1, a.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
        for r, S in ((0.2, 64), (0.4, 32)):
            assert torch.equal(_ext.ball_query(q, x, r, S), ref_ext.ball_query(q, x, r, S))
        d2a, ia = _ext.three_nn(x[:, :500].contiguous(), q)
        d2b, ib = ref_ext.three_nn(x[:, :500].contiguous(), q)
        assert torch.equal(ia, ib) and torch.equal(d2a, d2b)
        f = torch.randn(x.shape[0], 16, q.shape[1], device=DEV)
        w = torch.rand(x.shape[0], 500, 3, device=DEV)
        assert torch.equal(_ext.three_interpolate(f, ia, w), ref_ext.three_interpolate(f, ib, w))
        feats = torch.randn(x.shape[0], 8, x.shape[1], device=DEV)
        idx = ref_ext.ball_query(q, x, 0.2, 64)
        assert torch.equal(_ext.group_points(feats, idx), ref_ext.group_points(feats, idx))


def test_fused_group_vs_reference_python_sequence(ref_ext):
    """QueryAndGroup.forward op by op on the reference kernels + torch CUDA elementwise ops (pointnet2_utils.py:319-344),
    including torch's CUDA lowering of `grouped_xyz /= radius`."""
    cloud = cu(scannet_like_batch(1, 20000, seed0=2)[..., :3].copy())
    feats = torch.randn(1, 16, 20000, device=DEV)
    inds = ref_ext.furthest_point_sampling(cloud, 512)
    new_xyz = ref_ext.gather_points(cloud.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()
    for radius, S in ((0.2, 64), (0.3, 16)):
        idx = ref_ext.ball_query(new_xyz, cloud, radius, S)
        gx = ref_ext.group_points(cloud.transpose(1, 2).contiguous(), idx)
        gx -= new_xyz.transpose(1, 2).unsqueeze(-1)
        gx /= radius
        ref = torch.cat([gx, ref_ext.group_points(feats, idx)], dim=1)
        out, _, _ = pointnet2_utils.fused_query_and_group(cloud, new_xyz, feats, radius, S, True, True)
        assert torch.equal(out, ref)


@pytest.mark.parametrize("N,M,r,S", [(9000, 700, 0.15, 32), (20000, 256, 0.3, 64), (12000, 300, 0.05, 16), (30000, 64, 1.0, 128)])
def test_ball_query_grid_path_bit_exact(N, M, r, S):
    """N >= 8192 takes the uniform-grid candidate search; it must reproduce the brute-force semantics bit for bit:
    queries outside the cloud's bounding box, a dense blob with > 512 hits (brute-force fallback), duplicates."""
    rng = np.random.default_rng(N + S)
    cloud = rng.uniform(-2, 2, (2, N, 3)).astype(np.float32)
    cloud[:, :1500] = rng.normal(0, 0.02, (2, 1500, 3)).astype(np.float32) + 0.5   # dense blob
    cloud[:, 2000:2100] = cloud[:, 1900:2000]                                       # duplicates
    q = cloud[:, rng.choice(N, M, replace=False)].copy()
    q[:, 0] = (0.5, 0.5, 0.5)          # centre of the blob: thousands of hits
    q[:, 1] = (2.05, 0.0, 0.0)         # just outside the bounding box
    q[:, 2] = (40.0, 40.0, 40.0)       # far outside: no hits
    q[:, 3] = (-2.0, -2.0, -2.0)       # corner
    got = _ext.ball_query(cu(q), cu(cloud), r, S).cpu().numpy()
    assert np.array_equal(got, oracle.ball_query(q, cloud, r, S))
    if S <= 128:
        feats = rng.normal(size=(2, 6 if S != 32 else 70, N)).astype(np.float32)
        out, gxyz, idx = pointnet2_utils.fused_query_and_group(cu(cloud), cu(q), cu(feats), r, S, True, True,
                                                               ret_grouped_xyz=True, ret_idx=True)
        rf, rg, ri = model_ref.query_and_group(torch.from_numpy(cloud), torch.from_numpy(q), torch.from_numpy(feats),
                                               r, S, True, True, recip=True)
        assert np.array_equal(idx.cpu().numpy(), ri.numpy()) and torch.equal(out.cpu(), rf)


def test_fps_fused_new_xyz_output():
    """rfd_furthest_point_sampling_xyz: the coordinates emitted by the FPS kernel are exactly the gathered rows."""
    for cloud, m in ((tricky_cloud(4096, seed=1), 512), (scannet_like_batch(2, 30000, seed0=4)[..., :3].copy(), 700),
                     (np.zeros((1, 64, 3), np.float32), 5)):
        x = cu(cloud)
        idx, new_xyz = pointnet2_utils.fps_with_xyz(x, m)
        assert torch.equal(idx, _ext.furthest_point_sampling(x, m))
        assert torch.equal(new_xyz, torch.gather(x, 1, idx.long().unsqueeze(-1).expand(-1, -1, 3)))


def test_fps_prefix_check_and_conditional_sampler():
    """rfd_fps_prefix_check proves FPS(x)[:m] == arange(m) for FPS-ordered inputs; whatever the verdict, the conditional
    sampler returns exactly what the plain sampler returns (ties, duplicates, skipped points, unordered clouds)."""
    scene = cu(scannet_like_batch(3, 30000, seed0=11)[..., :3].copy())
    i1, x1 = pointnet2_utils.fps_with_xyz(scene, 2048)
    lattice = cu(np.random.default_rng(8).integers(-3, 4, (3, 2048, 3)).astype(np.float32))       # massive ties
    _, lat_fps = pointnet2_utils.fps_with_xyz(lattice, 2048)                                     # ... in FPS order
    tricky = cu(np.concatenate([tricky_cloud(2048, seed=1)] * 3))
    _, tricky_fps = pointnet2_utils.fps_with_xyz(tricky, 2048)
    unordered = cu(uniform_cloud(3, 2048, seed=5))
    mixed = torch.stack([x1[0], unordered[1], x1[2]]).contiguous()                               # per-scene verdicts
    from rfdnet_b200 import _lib
    lib = _lib.load()
    for name, x, expect in (("fps-ordered", x1, [1, 1, 1]), ("unordered", unordered, [0, 0, 0]), ("mixed", mixed, [1, 0, 1]),
                            ("lattice", lat_fps, None), ("tricky", tricky_fps, None)):
        for m in (1024, 2048, 1, 700):
            ref_i, ref_x = pointnet2_utils.fps_with_xyz(x, m)
            got_i, got_x = pointnet2_utils.fps_with_xyz(x, m, try_prefix=True)
            assert torch.equal(ref_i, got_i) and torch.equal(ref_x, got_x), (name, m)
            ws = torch.empty((3, m), device=DEV)
            flag = torch.full((3,), -1, dtype=torch.int32, device=DEV)
            _lib.check(lib.rfd_fps_prefix_check(x.data_ptr(), 3, x.shape[1], m, ws.data_ptr(), flag.data_ptr(), 0), "check")
            torch.cuda.synchronize()
            ident = (ref_i.cpu() == torch.arange(m, dtype=torch.int32)).all(dim=1).int().tolist()
            f = flag.cpu().tolist()
            assert all(fi <= ii for fi, ii in zip(f, ident)), (name, m, f, ident)   # never claims an identity that is not one
            if expect is not None and m > 1:
                assert f == expect, (name, m, f)
    # the backbone case: SA2 samples SA1's samples
    i2, _ = pointnet2_utils.fps_with_xyz(x1, 1024, try_prefix=True)
    assert torch.equal(i2.cpu(), torch.arange(1024, dtype=torch.int32).expand(3, -1))


def test_empty_and_degenerate_inputs():
    """zero-sized batches / queries / samples return empty tensors without launching; 1-point clouds work."""
    z3 = torch.zeros((0, 16, 3), device=DEV)
    assert _ext.furthest_point_sampling(z3, 4).shape == (0, 4)
    x = cu(uniform_cloud(2, 64, seed=1))
    assert _ext.furthest_point_sampling(x, 0).shape == (2, 0)
    assert _ext.ball_query(x[:, :0].contiguous(), x, 0.2, 8).shape == (2, 0, 8)
    assert _ext.ball_query(x[:, :5].contiguous(), x, 0.2, 0).shape == (2, 5, 0)
    f = torch.randn(2, 3, 64, device=DEV)
    assert _ext.group_points(f, torch.zeros((2, 0, 4), dtype=torch.int32, device=DEV)).shape == (2, 3, 0, 4)
    g = _ext.group_points_grad(torch.zeros((2, 3, 0, 4), device=DEV), torch.zeros((2, 0, 4), dtype=torch.int32, device=DEV), 64)
    assert g.shape == (2, 3, 64) and float(g.abs().sum()) == 0.0
    d2, ix = _ext.three_nn(x[:, :0].contiguous(), x)
    assert d2.shape == (2, 0, 3) and ix.shape == (2, 0, 3)
    one = cu(uniform_cloud(1, 1, seed=2))
    assert _ext.furthest_point_sampling(one, 3).cpu().tolist() == [[0, 0, 0]]
    assert _ext.ball_query(one, one, 0.1, 4).cpu().tolist() == [[[0, 0, 0, 0]]]
    # negative / zero radius: no hits -> zero rows (reference: d2 < r*r never true for r = 0)
    assert int(_ext.ball_query(x[:, :7].contiguous(), x, 0.0, 4).abs().sum()) == 0


def test_error_paths_raise():
    x = cu(uniform_cloud(1, 32, seed=1))
    with pytest.raises(RuntimeError, match="contiguous"):
        _ext.furthest_point_sampling(x.transpose(1, 2).transpose(1, 2)[:, ::2], 4)
    with pytest.raises(RuntimeError, match="float tensor"):
        _ext.furthest_point_sampling(x.double(), 4)
    with pytest.raises(RuntimeError, match="int tensor"):
        _ext.gather_points(torch.randn(1, 3, 32, device=DEV), torch.zeros((1, 4), dtype=torch.int64, device=DEV))
    with pytest.raises(RuntimeError, match="query_and_group"):
        pointnet2_utils.fused_query_and_group(x, x[:, :4].contiguous(), None, 0.2, 2000, True, True)  # nsample > 1024
    with pytest.raises(RuntimeError, match="float tensor"):
        pointnet2_utils.fused_query_and_group(x.double(), x[:, :4].contiguous(), None, 0.2, 8, True, True)
    with pytest.raises(RuntimeError, match="float tensor"):
        pointnet2_utils.fps_with_xyz(x.half(), 4)


def test_c_abi_device_round_trip(tmp_path):
    """tests/c/abi_check.c on the GPU box: a torch-free C99 process allocates with the CUDA runtime C API and calls FPS +
    fused query-and-group through the C ABI."""
    import os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "rfdnet_b200")
    exe = tmp_path / "abi_check"
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(root, "include"), "-isystem",
           "/usr/local/cuda/include", os.path.join(root, "tests", "c", "abi_check.c"), "-o", str(exe), "-L", libdir,
           "-lrfdnet_b200", "-L", "/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + libdir,
           "-Wl,-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "device round trip ok" in r.stdout, r.stdout + r.stderr
