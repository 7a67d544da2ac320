"""The drop-in boundary proved on hardware: the reference's UNMODIFIED pointnet2_utils.py and
pointnet2_modules.py (staged by oracle/build_ref.py under the git-ignored oracle/_ref/py/) are executed twice on the GPU --
once over rfdnet_b200._ext (the sm_100a C-ABI library) and once over the reference's own CUDA kernels recompiled for
sm_100 (oracle/_ref/_ref_ext.so) -- forward AND backward, and must agree."""
import numpy as np
import pytest
import torch

from oracle import build_ref
from rfdnet_b200 import _ext as ours_ext
from rfdnet_b200.synth import scannet_like_batch, seeded_fill

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def both():
    ref_ext = build_ref.load()
    if ref_ext is None:
        pytest.skip("oracle/_ref/_ref_ext.so not present")
    a = build_ref.load_py(ours_ext, "ours")
    b = build_ref.load_py(ref_ext, "ref")
    if a is None or b is None:
        pytest.skip("oracle/_ref/py not staged")
    return a, b


def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def test_reference_python_runs_on_the_dropin_ext(both):
    (u_o, m_o), (u_r, m_r) = both
    assert u_o._ext is ours_ext and u_r._ext is not ours_ext
    assert u_o.__file__ == u_r.__file__                      # the same unmodified reference file, two bindings


def test_sa_module_forward_backward_matches_reference_kernels(both):
    _no_tf32()
    (u_o, m_o), (u_r, m_r) = both
    pc = torch.from_numpy(scannet_like_batch(2, 6000, seed0=3)).to(DEV)
    xyz = pc[..., :3].contiguous()
    g = torch.Generator().manual_seed(0)
    feats0 = torch.randn(2, 8, 6000, generator=g).to(DEV)
    outs = []
    for mods in (m_o, m_r):
        sa = mods.PointnetSAModuleVotes(npoint=256, radius=0.3, nsample=16, mlp=[8, 32, 64], use_xyz=True,
                                        normalize_xyz=True)
        seeded_fill(sa, 7)
        sa = sa.to(DEV).train()
        feats = feats0.clone().requires_grad_(True)
        new_xyz, new_feats, inds = sa(xyz, feats)
        loss = (new_feats * torch.linspace(0.5, 1.5, new_feats.shape[-1], device=DEV)).sum()
        loss.backward()
        outs.append((new_xyz.detach(), new_feats.detach(), inds, feats.grad.clone(),
                     sa.mlp_module[0].weight.grad.clone()))
    (x_o, f_o, i_o, gf_o, gw_o), (x_r, f_r, i_r, gf_r, gw_r) = outs
    assert torch.equal(i_o, i_r) and torch.equal(x_o, x_r)            # FPS + gather: bit-exact
    assert torch.allclose(f_o, f_r, atol=1e-5, rtol=1e-5)             # same grouped tensor -> same cuDNN result
    assert torch.allclose(gf_o, gf_r, atol=1e-4, rtol=1e-4)           # scatter-add grads (atomics: order differs)
    assert torch.allclose(gw_o, gw_r, atol=1e-3, rtol=1e-4)


def test_fp_module_and_grouper_forward_backward_match_reference_kernels(both):
    _no_tf32()
    (u_o, m_o), (u_r, m_r) = both
    g = torch.Generator().manual_seed(1)
    unk = torch.rand(2, 500, 3, generator=g).to(DEV)
    kn = torch.rand(2, 120, 3, generator=g).to(DEV)
    uf0 = torch.randn(2, 12, 500, generator=g).to(DEV)
    kf0 = torch.randn(2, 20, 120, generator=g).to(DEV)
    outs = []
    for utils, mods in ((u_o, m_o), (u_r, m_r)):
        fp = mods.PointnetFPModule(mlp=[32, 24, 16])
        seeded_fill(fp, 9)
        fp = fp.to(DEV).train()
        uf, kf = uf0.clone().requires_grad_(True), kf0.clone().requires_grad_(True)
        out = fp(unk, kn, uf, kf)
        out.square().sum().backward()
        # the reference's own QueryAndGroup (ball_query + 2 x group_points + sub + div + cat) incl. backward
        qg = utils.QueryAndGroup(0.25, 16, use_xyz=True, normalize_xyz=True)
        kf2 = kf0.clone().requires_grad_(True)
        grouped = qg(kn, unk[:, :64].contiguous(), kf2)
        grouped.sum().backward()
        outs.append((out.detach(), uf.grad.clone(), kf.grad.clone(), grouped.detach(), kf2.grad.clone()))
    for a, b, tol in zip(outs[0], outs[1], (1e-5, 1e-4, 1e-4, 0.0, 1e-4)):
        if tol == 0.0:
            assert torch.equal(a, b)
        else:
            assert torch.allclose(a, b, atol=tol, rtol=1e-4)
