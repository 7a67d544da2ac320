"""rfd_extract_mesh (csrc/extract_mesh.cu) against the oracle restatement of Generator3D.extract_mesh
(generator.py:145-168 + PyMCubes 0.1.2): vertices bit-exact in f64 (same operation order), triangles identical."""
import numpy as np
import pytest
import torch

from mesh_checks import assert_closed_oriented
from oracle import cpu_ref
from rfdnet_b200 import generator, onet
from rfdnet_b200.synth import seeded_fill

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _fields(B, R, seed, smooth=True):
    rng = np.random.default_rng(seed)
    if not smooth:
        return rng.normal(size=(B, R, R, R)).astype(np.float32)
    ax = np.linspace(-1, 1, R)
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing="ij")
    out = []
    for b in range(B):
        f = np.full((R, R, R), -1.0)
        for _ in range(rng.integers(1, 4)):
            c, r = rng.uniform(-0.5, 0.5, 3), rng.uniform(0.2, 0.7)
            f = np.maximum(f, 2.0 * (r - np.sqrt((X - c[0]) ** 2 + (Y - c[1]) ** 2 + (Z - c[2]) ** 2)))
        out.append(f + 0.05 * rng.normal(size=f.shape))
    return np.stack(out).astype(np.float32)


def _check_against_oracle(fields, threshold=0.5, padding=0.1):
    B, R = fields.shape[0], fields.shape[1]
    lg = torch.from_numpy(fields.reshape(B, -1)).to(DEV)
    mb64 = generator.extract_meshes(lg, R, threshold, padding, vertex_dtype=torch.float64,
                                    vertices_per_object=3 * (R + 2) ** 3, triangles_per_object=5 * (R + 2) ** 3)
    mb32 = generator.extract_meshes(lg, R, threshold, padding, vertex_dtype=torch.float32,
                                    vertices_per_object=3 * (R + 2) ** 3, triangles_per_object=5 * (R + 2) ** 3)
    total_v = 0
    for b in range(B):
        v_ref, t_ref, keys = cpu_ref.extract_mesh(fields[b], threshold, padding)
        order = np.argsort(keys, kind="stable")           # oracle (PyMCubes creation order) -> (owner point, axis) order
        rank = np.empty_like(order)
        rank[order] = np.arange(len(order))
        v, t = mb64.mesh(b)
        assert v.dtype == np.float64 and v.shape == v_ref.shape and t.shape == t_ref.shape, (b, v.shape, v_ref.shape)
        assert np.array_equal(v, v_ref[order])            # bit-exact doubles
        assert np.array_equal(t, rank[t_ref] if len(t_ref) else t_ref)   # same triangles in the same (PyMCubes) order
        v32, t32 = mb32.mesh(b)
        assert v32.dtype == np.float32 and np.array_equal(v32, v.astype(np.float32)) and np.array_equal(t32, t)
        total_v += len(v)
    return mb64, total_v


def test_smooth_fields_match_oracle_bit_exact():
    mb, nv = _check_against_oracle(_fields(6, 32, 1))
    assert nv > 1000
    for b in range(6):
        assert_closed_oriented(mb.mesh(b)[1])


def test_noise_fields_every_case_match_oracle():
    _check_against_oracle(_fields(3, 12, 2, smooth=False))
    _check_against_oracle(_fields(2, 32, 3, smooth=False))       # ~50 k vertices per object: exercises the id range


@pytest.mark.parametrize("R", [1, 2, 3, 5, 31])
def test_small_and_odd_resolutions(R):
    _check_against_oracle(_fields(4, R, 10 + R, smooth=False))


def test_empty_full_and_threshold():
    f = np.stack([np.full((8, 8, 8), -3.0, np.float32), np.full((8, 8, 8), 4.0, np.float32),
                  np.zeros((8, 8, 8), np.float32)])
    mb, _ = _check_against_oracle(f)
    v, t, r = mb.to_host()
    assert r[0][1] == 0 and r[0][3] == 0 and r[2][1] == 0            # empty; value == threshold is outside
    assert r[1][1] == 6 * 64 and r[1][3] == 2 * r[1][1] - 4          # a closed box made by the -1e6 padding
    _check_against_oracle(_fields(2, 16, 4), threshold=0.3, padding=0.25)   # non-zero logit threshold, other box size


def test_pool_overflow_is_reported_not_written():
    f = _fields(4, 16, 5)
    lg = torch.from_numpy(f.reshape(4, -1)).to(DEV)
    mb = generator.extract_meshes(lg, 16, vertices_per_object=40, triangles_per_object=80)
    with pytest.raises(RuntimeError, match="did not fit"):
        mb.to_host()
    rng = mb.ranges.cpu().numpy()
    for b in range(4):
        v_ref, t_ref, _ = cpu_ref.extract_mesh(f[b])
        assert rng[b][1] == len(v_ref) and rng[b][3] == len(t_ref)    # counts are valid even when nothing was written
    with pytest.raises(RuntimeError):
        generator.extract_meshes(lg.cpu(), 16)
    with pytest.raises(ValueError):
        generator.extract_meshes(lg, 32)


def test_decoder_logits_to_meshes_full_size():
    """256 objects x 32^3 straight from the tensor-core decoder: every mesh closed + oriented, 8 checked against the oracle."""
    dec = onet.DecoderCBatchNorm(dim=3, z_dim=32, c_dim=512).eval()
    seeded_fill(dec, 31)
    dec = dec.to(DEV)
    grid = onet.make_3d_grid(32, 1.1, DEV)
    g = torch.Generator().manual_seed(3)
    c = torch.randn(256, 512, generator=g).to(DEV)
    z = torch.zeros(256, 32, device=DEV)
    with torch.no_grad():
        lg = dec.decode(grid, z, c)
    mb = generator.extract_meshes(lg, 32, vertices_per_object=3 * 34 ** 3 // 4, triangles_per_object=5 * 34 ** 3 // 4)
    v, t, r = mb.to_host()
    assert len(mb) == 256 and r[:, 1].sum() == len(v) and r[:, 3].sum() == len(t)
    # ranges tile the pools without overlap
    iv = sorted((int(a), int(n)) for a, n, _, _ in r)
    assert all(iv[i][0] + iv[i][1] == iv[i + 1][0] for i in range(len(iv) - 1)) and iv[0][0] == 0
    lgc = lg.cpu().numpy().reshape(256, 32, 32, 32)
    for b in range(0, 256, 32):
        v_ref, t_ref, keys = cpu_ref.extract_mesh(lgc[b])
        order = np.argsort(keys, kind="stable")
        vb, tb = mb.mesh(b)
        assert np.array_equal(vb, v_ref[order].astype(np.float32))
        rank = np.empty_like(order); rank[order] = np.arange(len(order))
        assert np.array_equal(tb, rank[t_ref] if len(t_ref) else t_ref)
    for b in range(256):
        vb, tb = mb.mesh(b)
        if len(tb):
            assert tb.max() == len(vb) - 1 and tb.min() == 0
            assert_closed_oriented(tb)
            assert np.abs(vb).max() <= 0.55 * (1 + 2.0 / 31) + 1e-6
    print(f"256 objects: {len(v)} vertices, {len(t)} triangles, {mb.d2h_bytes() / 1e6:.1f} MB vs logits {lg.numel() * 4 / 1e6:.1f} MB")


def test_run_host_mesh_result_equals_logits_path():
    """SceneHotPath.run_host(result='mesh') returns the meshes of exactly the logits run_host(result='logits') returns."""
    from rfdnet_b200.pipeline import SceneHotPath
    from rfdnet_b200.synth import scannet_like_batch
    net = SceneHotPath().eval()
    seeded_fill(net, 2)
    net = net.to(DEV)
    pc = torch.from_numpy(scannet_like_batch(1, 20000, seed0=3)).pin_memory()
    codes = torch.randn(256, 512, generator=torch.Generator().manual_seed(1)).pin_memory()
    logits_host = torch.empty((256, 32768)).pin_memory()
    net.run_host(pc, codes, logits_host, torch.device(DEV), result="logits")
    torch.cuda.synchronize()
    h2d, d2h = net.run_host(pc, codes, logits_host, torch.device(DEV), result="mesh", chunks=3)
    torch.cuda.synchronize()
    v, t, r = net.last_meshes
    assert d2h == v.nbytes + t.nbytes + r.nbytes + 24 * 3 + 4 * 256 * 2 and h2d == pc.numel() * 4 + codes.numel() * 4
    assert r[:, 1].sum() == len(v) and r[:, 3].sum() == len(t)
    lg = logits_host.numpy().reshape(256, 32, 32, 32)
    for b in (0, 100, 255):
        v_ref, t_ref, keys = cpu_ref.extract_mesh(lg[b])
        order = np.argsort(keys, kind="stable")
        rank = np.empty_like(order); rank[order] = np.arange(len(order))
        vo, nv, to, nt = (int(x) for x in r[b])
        assert np.array_equal(v[vo:vo + nv], v_ref[order].astype(np.float32))
        assert np.array_equal(t[to:to + nt], rank[t_ref] if len(t_ref) else t_ref)
    # a second step reuses the (now large enough) pinned pools and overlaps the copies: same result
    v0, t0, r0 = v.copy(), t.copy(), r.copy()
    net.run_host(pc, codes, logits_host, torch.device(DEV), result="mesh", chunks=3)
    torch.cuda.synchronize()
    v, t, r = net.last_meshes
    for b in (0, 17, 255):
        assert np.array_equal(v[r[b][0]:r[b][0] + r[b][1]], v0[r0[b][0]:r0[b][0] + r0[b][1]])
        assert np.array_equal(t[r[b][2]:r[b][2] + r[b][3]], t0[r0[b][2]:r0[b][2] + r0[b][3]])
    net.run_host(pc, codes, logits_host, torch.device(DEV), result="bits")
    torch.cuda.synchronize()
    bits = net.last_bits.numpy().view(np.uint32)
    occ = (lg.reshape(256, -1) >= 0.0)
    assert np.array_equal(np.unpackbits(bits.view(np.uint8), bitorder="little").reshape(256, -1).astype(bool), occ)
    with pytest.raises(RuntimeError, match="did not fit"):
        net.run_host(pc, codes, logits_host, torch.device(DEV), result="mesh", mesh_capacity=(8, 8))


def test_scene_generation_pipeline_composes_the_pieces():
    """pipeline.SceneGeneration = ISCNet.generate's device part (network.py:56-153): detection -> SkipPropagation ->
    decoder -> meshes; equals the pieces called one by one, for all proposals and for a caller-chosen subset."""
    from rfdnet_b200.pipeline import SceneGeneration
    from rfdnet_b200.synth import scannet_like_batch
    net = SceneGeneration().eval()
    seeded_fill(net, 6)
    net = net.to(DEV)
    pc = torch.from_numpy(scannet_like_batch(1, 20000, seed0=21)).to(DEV)
    ids = torch.tensor([[3, 200, 17, 255, 64, 9, 100, 31]], device=DEV)
    out = net(pc, proposal_ids=ids)
    assert out["codes"].shape == (8, 512) and out["logits"].shape == (8, 32768) and len(out["meshes"]) == 8
    with torch.no_grad():
        ep, pf = net.detection(pc, export_proposal_feature=True)
        ang = SceneGeneration.heading_angles(ep, 12)
        assert float(ang.max()) <= np.pi + 1e-6 and float(ang.min()) > -np.pi - 1e-6
        codes = net.skip_propagation.generate(ep["center"][:, ids[0]].contiguous(), ang[:, ids[0]].contiguous(),
                                              pf[:, :, ids[0]].contiguous(), pc)
    assert torch.equal(out["codes"], codes.transpose(1, 2).reshape(8, 512))
    lg = out["logits"].cpu().numpy().reshape(8, 32, 32, 32)
    for b in (0, 7):
        v_ref, t_ref, keys = cpu_ref.extract_mesh(lg[b])
        order = np.argsort(keys, kind="stable")
        v, t = out["meshes"].mesh(b)
        assert np.array_equal(v, v_ref[order].astype(np.float32)) and len(t) == len(t_ref)
    full = net(pc, meshes=False)
    assert full["codes"].shape == (256, 512) and full["meshes"] is None and bool(torch.isfinite(full["logits"]).all())
    assert torch.allclose(full["codes"][ids[0]], out["codes"], atol=1e-5)
