"""Generate the golden fixtures under tests/golden/ by running the REAL reference Python modules
(/root/reference, PUBLIC UNTRUSTED CONTENT: imported for behaviour only) on CPU.

Run in the build container only (`python tests/golden/make_golden.py`); the GPU box has no /root/reference
and only reads the committed .npz files.

How the reference is imported: `models` is stubbed so that models/__init__.py -> models/loss.py
(which JIT-builds a CUDA extension at import) never runs; `pointnet2_ops._ext` is provided by the CPU oracle
adapter oracle/torch_ext.py, so pointnet2_utils.py / pointnet2_modules.py run unmodified on CPU.  The index
kernels underneath are therefore the oracle's, while every line of module glue, conv/BN/pool and the whole
ONet decoder is the reference's own code.  Each fixture is also cross-checked here against
oracle/model_ref.py so that the restatement is pinned to the reference.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import model_ref, torch_ext  # noqa: E402
from rfdnet_b200.synth import scannet_like_batch, seeded_fill, tricky_cloud, uniform_cloud  # noqa: E402


def import_reference():
    sys.path.insert(0, REF)
    pkg = types.ModuleType("pointnet2_ops")
    pkg.__path__ = [os.path.join(REF, "external/pointnet2_ops_lib/pointnet2_ops")]
    pkg._ext = torch_ext
    sys.modules["pointnet2_ops"] = pkg
    sys.modules["pointnet2_ops._ext"] = torch_ext
    for name, path in [("models", "models"), ("models.iscnet", "models/iscnet"),
                       ("models.iscnet.modules", "models/iscnet/modules")]:
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, path)]
        sys.modules[name] = m
    reg = types.ModuleType("models.registers")

    class _Reg:
        def register_module(self, cls):
            return cls
    reg.MODULES = _Reg()
    reg.METHODS = _Reg()
    reg.LOSSES = _Reg()
    sys.modules["models.registers"] = reg
    mods = {}
    mods["p2m"] = importlib.import_module("external.pointnet2_ops_lib.pointnet2_ops.pointnet2_modules")
    mods["backbone"] = importlib.import_module("models.iscnet.modules.pointnet2backbone")
    mods["vote"] = importlib.import_module("models.iscnet.modules.vote_module")
    mods["prop"] = importlib.import_module("models.iscnet.modules.proposal_module")
    mods["dec"] = importlib.import_module("models.iscnet.modules.occ_decoder")
    # external/common.py:3 imports the (unbuilt, unrelated) kd-tree extension; stub that one symbol
    kd = types.ModuleType("external.libkdtree.pykdtree.kdtree")
    kd.KDTree = None
    for name in ("external.libkdtree", "external.libkdtree.pykdtree"):
        stub = types.ModuleType(name)
        stub.__path__ = []
        sys.modules[name] = stub
    sys.modules["external.libkdtree.pykdtree.kdtree"] = kd
    mods["common"] = importlib.import_module("external.common")
    return mods


class _Cfg:
    def __init__(self):
        self.config = {"data": {"use_color_detection": False, "no_height": False, "vote_factor": 1,
                                "num_target": 256, "cluster_sampling": "vote_fps"}}

        class DC:
            num_class, num_heading_bin, num_size_cluster = 8, 12, 8
            mean_size_arr = np.ones((8, 3), np.float32)
        self.dataset_config = DC()


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    M = import_reference()
    out = {}

    # ---- G1: index-level known answers on config 1 (4096 pts, npoint 512, r 0.2 / 0.4, nsample 32)
    for tag, cloud in (("uniform", uniform_cloud(2, 4096, seed=0)), ("tricky", tricky_cloud(4096, seed=1))):
        import oracle
        fps = oracle.furthest_point_sampling(cloud, 512)
        new_xyz = np.take_along_axis(cloud, fps[..., None].astype(np.int64).repeat(3, -1), 1)
        bq = oracle.ball_query(new_xyz, cloud, 0.2, 32)
        bq4 = oracle.ball_query(new_xyz, cloud, 0.4, 32)
        d2, nn = oracle.three_nn(cloud[:, :700], new_xyz)
        out[f"c1_{tag}_fps"] = fps
        out[f"c1_{tag}_bq02"] = bq.astype(np.int16)
        out[f"c1_{tag}_bq04"] = bq4.astype(np.int16)
        out[f"c1_{tag}_nn_idx"] = nn.astype(np.int16)
        out[f"c1_{tag}_nn_d2"] = d2

    # ---- G2: one SA module through the reference's own PointnetSAModuleVotes (CPU => true division by radius)
    sa = M["p2m"].PointnetSAModuleVotes(npoint=128, radius=0.3, nsample=16, mlp=[5, 32, 32, 64], use_xyz=True,
                                        normalize_xyz=True).eval()
    seeded_fill(sa, 11)
    g = torch.Generator().manual_seed(3)
    xyz = torch.from_numpy(uniform_cloud(2, 1024, seed=5))
    feats = torch.randn(2, 5, 1024, generator=g)
    with torch.no_grad():
        nx, nf, ind = sa(xyz, feats)
    sd = {"sa." + k: v for k, v in sa.state_dict().items()}
    rx, rf, ri = model_ref.sa_module(xyz, feats, sd, "sa", 128, 0.3, 16, recip=False)
    assert torch.equal(ri, ind) and torch.equal(rx, nx), "model_ref.sa_module index mismatch"
    assert torch.allclose(rf, nf, atol=1e-6, rtol=1e-6), float((rf - nf).abs().max())
    out.update(sa_new_xyz=nx.numpy(), sa_new_features=nf.numpy(), sa_inds=ind.numpy())

    # ---- G3: FP module
    fp = M["p2m"].PointnetFPModule(mlp=[64 + 16, 64, 32]).eval()
    seeded_fill(fp, 12)
    unk = torch.from_numpy(uniform_cloud(2, 300, seed=6))
    kn = torch.from_numpy(uniform_cloud(2, 64, seed=7))
    uf, kf = torch.randn(2, 16, 300, generator=g), torch.randn(2, 64, 64, generator=g)
    with torch.no_grad():
        fo = fp(unk, kn, uf, kf)
    sd = {"fp." + k: v for k, v in fp.state_dict().items()}
    ro = model_ref.fp_module(unk, kn, uf, kf, sd, "fp")
    assert torch.allclose(ro, fo, atol=1e-6, rtol=1e-6), float((ro - fo).abs().max())
    out.update(fp_out=fo.numpy())

    # ---- G4: backbone + voting + proposal on one 20000-point ScanNet-like scene (reference modules, eval)
    cfg = _Cfg()
    bb = M["backbone"].Pointnet2Backbone(cfg).eval()
    vm = M["vote"].VotingModule(cfg).eval()
    pm = M["prop"].ProposalModule(cfg).eval()
    seeded_fill(bb, 21)
    seeded_fill(vm, 22)
    seeded_fill(pm, 23)
    pc = torch.from_numpy(scannet_like_batch(1, 20000, seed0=100))
    with torch.no_grad():
        ep = bb(pc, {})
        vx, vf = vm(ep["fp2_xyz"], ep["fp2_features"])
        vf = vf.div(torch.norm(vf, p=2, dim=1).unsqueeze(1))
        ep["seed_xyz"] = ep["fp2_xyz"]
        ep2, _ = pm(vx, vf, dict(ep))
    sd = {}
    sd.update({"backbone." + k: v for k, v in bb.state_dict().items()})
    sd.update({"voting." + k: v for k, v in vm.state_dict().items()})
    sd.update({"detection." + k: v for k, v in pm.state_dict().items()})
    rep = model_ref.backbone(pc, sd, recip=False)
    for k in ("sa1_inds", "sa2_inds"):
        assert torch.equal(rep[k], ep[k]), k
    assert torch.allclose(rep["fp2_features"], ep["fp2_features"], atol=2e-5, rtol=1e-5)
    rvx, rvf = model_ref.voting(rep["fp2_xyz"], rep["fp2_features"], sd)
    assert torch.allclose(rvf, vf, atol=2e-5, rtol=1e-5)
    rax, rinds, rnet = model_ref.proposal(rvx, rvf, sd, recip=False)
    assert torch.equal(rinds, ep2["aggregated_vote_inds"])
    assert torch.allclose(rnet.transpose(2, 1)[:, :, 0:2], ep2["objectness_scores"], atol=1e-4, rtol=1e-4)
    out.update(det_sa1_inds=ep["sa1_inds"].numpy().astype(np.int32),
               det_sa2_inds_is_arange=np.array(torch.equal(ep["sa2_inds"][0].long(), torch.arange(1024))),
               det_sa4_features=ep["sa4_features"].numpy()[:, :, :32],
               det_fp2_features=ep["fp2_features"].numpy()[:, :, :64],
               det_vote_xyz=vx.numpy()[:, :128], det_agg_inds=ep2["aggregated_vote_inds"].numpy().astype(np.int32),
               det_objectness=ep2["objectness_scores"].numpy(), det_center=ep2["center"].numpy(),
               det_sem_cls=ep2["sem_cls_scores"].numpy())

    # ---- G5: ONet decoder (reference DecoderCBatchNorm, eval, every zero-initialised tensor re-randomised)
    dec = M["dec"].DecoderCBatchNorm(dim=3, z_dim=32, c_dim=512, hidden_size=256, n_blocks=5).eval()
    seeded_fill(dec, 31)
    grid = 1.1 * M["common"].make_3d_grid((-0.5,) * 3, (0.5,) * 3, (32,) * 3)
    assert torch.equal(grid, model_ref.make_3d_grid(32, 1.1))
    out["grid32_axis"] = (1.1 * torch.linspace(-0.5, 0.5, 32)).numpy()
    c = torch.randn(3, 512, generator=g)
    z = torch.zeros(3, 32)
    sel = torch.arange(0, 32768, 61)[:512]
    p = grid[sel].unsqueeze(0).expand(3, -1, -1).contiguous()
    with torch.no_grad():
        logits = dec(p, z, c)
        z2 = torch.randn(3, 32, generator=g) * 0.3
        logits_z = dec(p, z2, c)
    sd = dict(dec.state_dict())
    rl = model_ref.decoder(p, z, c, sd)
    assert torch.allclose(rl, logits, atol=1e-5, rtol=1e-5), float((rl - logits).abs().max())
    out.update(dec_c=c.numpy(), dec_sel=sel.numpy().astype(np.int32), dec_logits=logits.numpy(),
               dec_z2=z2.numpy(), dec_logits_z=logits_z.numpy())

    # ---- G6 (section 8f rank 1): STN_Group, the per-proposal grouping of SkipPropagation (r = 1.0, nsample = 1024)
    stn = M["p2m"].STN_Group(radius=1.0, nsample=1024, use_xyz=False, normalize_xyz=False).eval()
    seeded_fill(stn, 41, scale=0.3)
    pcs = torch.from_numpy(scannet_like_batch(1, 12000, seed0=55))
    sxyz = pcs[..., :3].contiguous()
    sfeat = torch.cat([pcs[..., 3:].transpose(1, 2), torch.randint(0, 5, (1, 1, 12000), generator=g).float()], dim=1).contiguous()
    box_xyz = sxyz[:, torch.tensor([5, 900, 4000, 7777, 11000])].contiguous() + 0.05
    orient = torch.tensor([[0.0, 0.7, -1.2, 2.5, 3.0]])
    with torch.no_grad():
        gx, gf = stn(sxyz, sfeat, box_xyz, orient)
    out.update(stn_grouped_xyz=gx.numpy()[:, :, :, ::8], stn_grouped_feat_sum=gf.sum(-1).numpy(),
               stn_feat_first=gf.numpy()[:, :, :, :4])
    out["keys_stn_group"] = np.array([f"{k}:{tuple(v.shape)}" for k, v in stn.state_dict().items()])

    # ---- state_dict key/shape inventories of the reference modules (checkpoint compatibility of the mirrors)
    for name, mod in (("backbone", bb), ("voting", vm), ("detection", pm), ("decoder", dec)):
        sdm = mod.state_dict()
        out[f"keys_{name}"] = np.array([f"{k}:{tuple(v.shape)}" for k, v in sdm.items()])

    np.savez_compressed(os.path.join(HERE, "golden.npz"), **out)
    sz = os.path.getsize(os.path.join(HERE, "golden.npz"))
    print("wrote golden.npz", sz, "bytes;", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
