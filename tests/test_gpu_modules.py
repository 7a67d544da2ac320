"""GPU tests of the module-level hot path (SA / FP / voting / proposal mirrors on the sm_100a kernels) against
oracle/model_ref (CPU) and the golden fixtures produced by the reference's own Python modules."""
import numpy as np
import pytest
import torch

from oracle import model_ref
from rfdnet_b200 import detection, pointnet2_modules
from rfdnet_b200.synth import scannet_like_batch, seeded_fill, uniform_cloud

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_sa_module_fused_vs_oracle_and_golden(golden):
    sa = pointnet2_modules.PointnetSAModuleVotes(npoint=128, radius=0.3, nsample=16, mlp=[5, 32, 32, 64],
                                                 use_xyz=True, normalize_xyz=True).eval()
    seeded_fill(sa, 11)
    g = torch.Generator().manual_seed(3)
    xyz = torch.from_numpy(uniform_cloud(2, 1024, seed=5))
    feats = torch.randn(2, 5, 1024, generator=g)
    sd = {"sa." + k: v for k, v in sa.state_dict().items()}
    rx, rf, ri = model_ref.sa_module(xyz, feats, sd, "sa", 128, 0.3, 16, recip=True)
    sa = sa.to(DEV)
    with torch.no_grad():
        nx, nf, ind = sa(xyz.to(DEV), feats.to(DEV))
    assert np.array_equal(ind.cpu().numpy(), golden["sa_inds"]) and torch.equal(ind.cpu(), ri)
    assert np.array_equal(nx.cpu().numpy(), golden["sa_new_xyz"])
    assert torch.allclose(nf.cpu(), rf, atol=1e-4, rtol=1e-4)
    assert np.allclose(nf.cpu().numpy(), golden["sa_new_features"], atol=1e-4, rtol=1e-4)
    # training-mode path = reference sequence on the drop-in ops + cuDNN (TF32 off: torch's default would be the
    # less accurate side): same numbers as the fused fp32 kernels
    tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        sa_ref_path = sa._forward_reference(xyz.to(DEV), feats.to(DEV), None)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    assert torch.equal(sa_ref_path[2], ind)
    assert torch.allclose(sa_ref_path[1], nf, atol=1e-4, rtol=1e-4)


def test_fp_module_vs_golden(golden):
    fp = pointnet2_modules.PointnetFPModule(mlp=[64 + 16, 64, 32]).eval()
    seeded_fill(fp, 12)
    g = torch.Generator().manual_seed(3)
    _ = torch.randn(2, 5, 1024, generator=g)  # keep the generator in step with make_golden.py
    unk = torch.from_numpy(uniform_cloud(2, 300, seed=6)).to(DEV)
    kn = torch.from_numpy(uniform_cloud(2, 64, seed=7)).to(DEV)
    uf, kf = torch.randn(2, 16, 300, generator=g).to(DEV), torch.randn(2, 64, 64, generator=g).to(DEV)
    fp = fp.to(DEV)
    with torch.no_grad():
        out = fp(unk, kn, uf, kf)
    assert np.allclose(out.cpu().numpy(), golden["fp_out"], atol=1e-4, rtol=1e-4)


def test_detection_hot_path_vs_golden(golden):
    net = detection.DetectionHotPath(1, 256).eval()
    seeded_fill(net.backbone, 21)
    seeded_fill(net.voting, 22)
    seeded_fill(net.detection, 23)
    net = net.to(DEV)
    pc = torch.from_numpy(scannet_like_batch(1, 20000, seed0=100)).to(DEV)
    with torch.no_grad():
        ep, _ = net(pc)
    assert np.array_equal(ep["sa1_inds"].cpu().numpy(), golden["det_sa1_inds"])
    assert np.array_equal(ep["aggregated_vote_inds"].cpu().numpy(), golden["det_agg_inds"])
    # golden was produced on CPU (true division by the radius), the GPU path multiplies by 1/r like torch-CUDA:
    # features agree to fp32 round-off
    assert np.allclose(ep["sa4_features"].cpu().numpy()[:, :, :32], golden["det_sa4_features"], atol=1e-4, rtol=1e-4)
    assert np.allclose(ep["fp2_features"].cpu().numpy()[:, :, :64], golden["det_fp2_features"], atol=1e-4, rtol=1e-4)
    assert np.allclose(ep["vote_xyz"].cpu().numpy()[:, :128], golden["det_vote_xyz"], atol=1e-4, rtol=1e-4)
    assert np.allclose(ep["objectness_scores"].cpu().numpy(), golden["det_objectness"], atol=2e-4, rtol=1e-3)
    assert np.allclose(ep["center"].cpu().numpy(), golden["det_center"], atol=2e-4, rtol=1e-3)
    assert np.allclose(ep["sem_cls_scores"].cpu().numpy(), golden["det_sem_cls"], atol=2e-4, rtol=1e-3)


def test_detection_80k_batch_shapes_and_determinism():
    net = detection.DetectionHotPath(1, 256).eval()
    seeded_fill(net, 5)
    net = net.to(DEV)
    pc = torch.from_numpy(scannet_like_batch(2, 80000, seed0=7)).to(DEV)
    with torch.no_grad():
        ep, _ = net(pc)
        ep2, _ = net(pc)
    assert ep["sa1_xyz"].shape == (2, 2048, 3) and ep["fp2_features"].shape == (2, 256, 1024)
    assert ep["objectness_scores"].shape == (2, 256, 2) and ep["center"].shape == (2, 256, 3)
    assert torch.equal(ep["sa2_inds"][0].long().cpu(), torch.arange(1024))
    for k in ("sa1_inds", "fp2_features", "center", "sem_cls_scores"):
        assert torch.equal(ep[k], ep2[k]), k  # no atomics on the forward path: bitwise reproducible
    assert torch.isfinite(ep["center"]).all()


CHAIN_TOL = {"x3": 1e-4, "fp16": 3e-3, "bf16": 2e-2}   # x max(1, output scale); x3 is the BASELINE config-2 bound


@pytest.mark.parametrize("mode", ["x3", "fp16", "bf16"])
@pytest.mark.parametrize("Ct,C1,C2,C3,M,S", [(4, 64, 64, 128, 2048, 64), (131, 128, 128, 256, 1024, 32),
                                             (259, 128, 128, 256, 512, 16), (259, 128, 128, 128, 100, 16),
                                             (20, 64, 128, 192, 37, 32), (70, 32, 48, 100, 9, 64)])
def test_chain_mlp_pooled_vs_fp32(Ct, C1, C2, C3, M, S, mode):
    """tcgen05 chain kernel on a materialised grouped tensor (dense rows + max over S) against the exact fp32
    CUDA-core path and against the reference's own module sequence in torch fp32."""
    from rfdnet_b200 import mlp
    seq = pointnet2_modules.build_shared_mlp([Ct, C1, C2, C3]).eval()
    seeded_fill(seq, Ct + C3)
    seq = seq.to(DEV)
    g = torch.Generator().manual_seed(M)
    x = torch.randn(2, Ct, M, S, generator=g).to(DEV)
    layers = mlp.fold_sequential(seq)
    ref = mlp.run_mlp(x.view(2, Ct, M * S), layers, pool_last=S)
    tc = mlp.ChainMlp(layers, xyz=0, mode=mode)
    assert tc.ok
    out, out_pm = tc.dense(x.view(2, Ct, M * S), pool=S, want_cm=True, want_pm=True)
    assert out.shape == (2, C3, M) and torch.equal(out_pm, out.transpose(1, 2))
    scale = float(ref.abs().max())
    err = float((out - ref).abs().max())
    print(f"chain {mode} Ct={Ct} S={S}: max|err| {err:.3e} scale {scale:.3f}")
    assert err <= CHAIN_TOL[mode] * max(1.0, scale)
    if mode == "x3":
        tf32 = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        try:
            with torch.no_grad():
                t = torch.nn.functional.max_pool2d(seq(x), kernel_size=[1, S]).squeeze(-1)
        finally:
            torch.backends.cudnn.allow_tf32 = tf32
        assert torch.allclose(out, t, atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("mode", ["x3", "fp16"])
@pytest.mark.parametrize("spec,L,relu_last", [([512, 256, 256], 1024, True), ([256, 256, 256, 259], 1024, False),
                                              ([128, 128, 128, 69], 256, False), ([80, 64, 32], 300, True),
                                              ([40, 24], 77, False), ([200, 512], 130, True)])
def test_chain_mlp_dense_heads_vs_fp32(spec, L, relu_last, mode):
    """FP-module, voting and proposal-head shapes (no pooling; last layer up to 512 wide in two column blocks;
    K0 = 512 streams through the resident A panels in two rounds)."""
    from rfdnet_b200 import mlp
    g = torch.Generator().manual_seed(L)
    layers = []
    for i in range(1, len(spec)):
        W = (torch.randn(spec[i], spec[i - 1], generator=g) / spec[i - 1] ** 0.5).to(DEV)
        s = (torch.rand(spec[i], generator=g) + 0.5).to(DEV)
        t = (torch.randn(spec[i], generator=g) * 0.2).to(DEV)
        layers.append((W, s, t, True if i < len(spec) - 1 else relu_last))
    x = torch.randn(2, spec[0], L, generator=g).to(DEV)
    ref = mlp.run_mlp(x, layers)
    tc = mlp.ChainMlp(layers, xyz=0, mode=mode)
    assert tc.ok
    out, out_pm = tc.dense(x, want_cm=True, want_pm=True)
    assert out.shape == ref.shape and torch.equal(out_pm, out.transpose(1, 2))
    err, scale = float((out - ref).abs().max()), float(ref.abs().max())
    print(f"chain dense {mode} {spec}: max|err| {err:.3e} scale {scale:.3f}")
    assert err <= CHAIN_TOL[mode] * max(1.0, scale)


def test_detection_16bit_paths_close_to_exact():
    net = detection.DetectionHotPath(1, 256).eval()
    seeded_fill(net, 5)
    net = net.to(DEV)
    pc = torch.from_numpy(scannet_like_batch(1, 80000, seed0=9)).to(DEV)

    def set_precision(p):
        for m in net.modules():
            if hasattr(m, "precision"):
                m.precision = p

    with torch.no_grad():
        set_precision('cuda')
        ep_c, _ = net(pc)
        set_precision('x3')
        ep, _ = net(pc)
        set_precision('fp16')
        ep_h, _ = net(pc)
        set_precision('bf16')
        ep_b, _ = net(pc)
    assert torch.equal(ep["sa1_inds"], ep_b["sa1_inds"]) and torch.equal(ep["sa1_inds"], ep_c["sa1_inds"])
    for k in ("sa1_features", "sa2_features", "sa4_features", "fp2_features"):
        f = ep_c[k]
        sc = max(1.0, float(f.abs().max()))
        assert float((f - ep[k]).abs().max()) <= 1e-4 * sc, k        # tensor cores, fp32-grade
        assert float((f - ep_h[k]).abs().max()) <= 1e-2 * sc, k
        assert float((f - ep_b[k]).abs().max()) <= 6e-2 * sc, k


def test_stn_group_vs_reference_golden(golden):
    """STN_Group (ball query r = 1.0, nsample = 1024 over the cloud, rotation, STN3d) on the
    sm_100a kernels against the fixture produced by the reference's own STN_Group on CPU."""
    from rfdnet_b200 import stn_group
    stn = stn_group.STN_Group(radius=1.0, nsample=1024, use_xyz=False, normalize_xyz=False).eval()
    seeded_fill(stn, 41, scale=0.3)
    stn = stn.to(DEV)
    g = torch.Generator().manual_seed(3)
    # keep the generator in step with tests/golden/make_golden.py (G2, G3, G5 draws precede G6)
    torch.randn(2, 5, 1024, generator=g); torch.randn(2, 16, 300, generator=g); torch.randn(2, 64, 64, generator=g)
    torch.randn(3, 512, generator=g); torch.randn(3, 32, generator=g)
    pcs = torch.from_numpy(scannet_like_batch(1, 12000, seed0=55))
    sxyz = pcs[..., :3].contiguous()
    sfeat = torch.cat([pcs[..., 3:].transpose(1, 2), torch.randint(0, 5, (1, 1, 12000), generator=g).float()], dim=1).contiguous()
    box_xyz = sxyz[:, torch.tensor([5, 900, 4000, 7777, 11000])].contiguous() + 0.05
    orient = torch.tensor([[0.0, 0.7, -1.2, 2.5, 3.0]])
    with torch.no_grad():
        gx, gf = stn(sxyz.to(DEV), sfeat.to(DEV), box_xyz.to(DEV), orient.to(DEV))
    assert gx.shape == (1, 3, 5, 1024) and gf.shape == (1, 2, 5, 1024)
    assert np.array_equal(gf.cpu().numpy()[:, :, :, :4], golden["stn_feat_first"])           # gathered rows: exact
    assert np.allclose(gf.sum(-1).cpu().numpy(), golden["stn_grouped_feat_sum"], rtol=1e-5, atol=1e-3)
    assert np.allclose(gx.cpu().numpy()[:, :, :, ::8], golden["stn_grouped_xyz"], atol=2e-4, rtol=1e-4)


@pytest.mark.parametrize("N,npoint,radius,S,C,mlp", [(4096, 512, 0.3, 32, 5, [5, 64, 64, 128]), (2048, 1024, 0.4, 32, 128, [128, 128, 128, 256]),
                                                     (1024, 256, 0.3, 16, 256, [256, 128, 128, 128]), (20000, 2048, 0.2, 64, 1, [1, 64, 64, 128]),
                                                     (3000, 100, 0.5, 128, 0, [0, 32, 64])])
def test_full_sa_fusion_vs_materialised_fp32_path(N, npoint, radius, S, C, mlp):
    """FPS -> ball query -> ONE kernel (gather + centre/normalise + tcgen05 MLP + max) WITHOUT
    materialising the grouped tensor, against the materialised fp32 CUDA-core path: same indices, features to 1e-4."""
    sa = pointnet2_modules.PointnetSAModuleVotes(npoint=npoint, radius=radius, nsample=S, mlp=list(mlp), use_xyz=True,
                                                 normalize_xyz=True, precision='cuda').eval()
    seeded_fill(sa, N + S)
    sa = sa.to(DEV)
    xyz = torch.from_numpy(scannet_like_batch(2, N, seed0=N)[..., :3].copy()).to(DEV)
    g = torch.Generator().manual_seed(C)
    feats = torch.randn(2, C, N, generator=g).to(DEV) if C else None
    with torch.no_grad():
        x1, f1, i1 = sa(xyz, feats)
        sa.precision = 'x3'
        x2, f2, i2, p2 = sa._forward_fused(xyz, feats, None, want_pm=True)
    assert torch.equal(i1, i2) and torch.equal(x1, x2)
    assert torch.equal(p2, f2.transpose(1, 2))
    err, scale = float((f1 - f2).abs().max()), float(f1.abs().max())
    print(f"fused SA N={N} S={S} C={C}: max|err| {err:.3e} scale {scale:.3f}")
    assert err <= 1e-4 * max(1.0, scale)


def test_stn_group_fused_path_vs_op_by_op_80k():
    """The four-launch inference path of STN_Group (rotated grouping kernel, two tcgen05 chains, affine kernel) against
    the module's own op-by-op autograd path (drop-in ball query / grouping + torch layers) on 80k-point scenes."""
    from rfdnet_b200 import stn_group
    _no_tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        stn = stn_group.STN_Group(radius=1.0, nsample=1024, use_xyz=False, normalize_xyz=True).eval()
        seeded_fill(stn, 43, scale=0.3)
        stn = stn.to(DEV)
        pcs = torch.from_numpy(scannet_like_batch(2, 80000, seed0=61)).to(DEV)
        xyz = pcs[..., :3].contiguous()
        g = torch.Generator().manual_seed(8)
        feats = torch.cat([pcs[..., 3:].transpose(1, 2), torch.randint(0, 9, (2, 1, 80000), generator=g).float().to(DEV)], 1).contiguous()
        sel = torch.randint(0, 80000, (16,), generator=g)
        box_xyz = (xyz[:, sel] + 0.1).contiguous()
        heading = (torch.rand(2, 16, generator=g) * 6.28 - 3.14).to(DEV)
        with torch.no_grad():
            gx, gf = stn(xyz, feats, box_xyz, heading)                      # fused
        with torch.enable_grad():
            rx, rf = stn(xyz, feats, box_xyz, heading)                      # op by op (grad mode switches the path)
        assert gx.shape == (2, 3, 16, 1024) and gf.shape == (2, 2, 16, 1024)
        assert torch.equal(gf, rf.detach())                                 # gathered rows: exact
        err = float((gx - rx.detach()).abs().max())
        print(f"STN_Group fused vs op-by-op: max|err| {err:.2e} (coordinate scale {float(rx.abs().max()):.2f})")
        assert err <= 2e-4
        # use_xyz=True variant: the xyz channels of the features stay UNROTATED, as in the reference
        stn2 = stn_group.STN_Group(radius=0.5, nsample=128, use_xyz=True, normalize_xyz=False).eval()
        seeded_fill(stn2, 44, scale=0.3)
        stn2 = stn2.to(DEV)
        with torch.no_grad():
            gx2, gf2 = stn2(xyz, feats, box_xyz, heading)
        with torch.enable_grad():
            rx2, rf2 = stn2(xyz, feats, box_xyz, heading)
        assert gf2.shape == (2, 5, 16, 128) and torch.equal(gf2, rf2.detach())
        assert float((gx2 - rx2.detach()).abs().max()) <= 2e-4
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = _no_tf32


@pytest.mark.parametrize("mode,tol", [("x3", 2e-4), ("fp16", 3e-2)])
def test_skip_propagation_generate_tensor_core_vs_torch(mode, tol):
    """SkipPropagation.generate: PointSeg + ResnetPointnet on the tcgen05 chain kernel (per-cloud biases instead of
    repeated global features, fused shortcut, epilogue max-pools) against the mirror's torch layers (fp32, TF32 off)."""
    from rfdnet_b200 import completion, completion_fast
    tf = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        sp = completion.SkipPropagation(input_feature_dim=1, c_dim=512, hidden_dim=512).eval()
        seeded_fill(sp, 23)
        sp = sp.to(DEV)
        pc = torch.from_numpy(scannet_like_batch(2, 30000, seed0=77)).to(DEV)
        g = torch.Generator().manual_seed(4)
        sel = torch.randint(0, 30000, (6,), generator=g)
        box_xyz = (pc[:, sel, :3] + 0.05).contiguous()
        heading = (torch.rand(2, 6, generator=g) * 6.28).to(DEV)
        box_feat = torch.randn(2, 128, 6, generator=g).to(DEV)
        with torch.no_grad():
            # balance the two classes of the seeded segmentation head, so that about half of the points survive the mask
            xyz, feats = sp._break_up_pc(pc)
            feats = torch.cat([feats, torch.zeros_like(feats)], dim=1)
            gx, gf = sp.stn(xyz, feats, box_xyz, heading)
            pts = torch.cat([gx, gf[:, :1]], dim=1).permute(0, 2, 3, 1).contiguous().view(12, 1024, -1)
            x, _, _ = sp.point_seg.feat(pts.transpose(1, 2).contiguous())
            ps = sp.point_seg
            for conv, bn in ((ps.conv1, ps.bn1), (ps.conv2, ps.bn2), (ps.conv3, ps.bn3)):
                x = torch.relu(bn(conv(x)))
            lg = ps.conv4(x)
            ps.conv4.bias[1] += torch.median(lg[:, 0] - lg[:, 1])
            sp.fast_precision = None
            ref = sp.generate(box_xyz, heading, box_feat, pc)
            sp.fast_precision = mode
            out = sp.generate(box_xyz, heading, box_feat, pc)
            # the masks of both paths (a flipped near-tie moves a whole point in or out of the encoder's input)
            _, mask = completion_fast.encode(sp, gx, gf, box_feat, mode)
            seg, _ = sp.point_seg(pts.transpose(1, 2).contiguous())
            mask_ref = torch.argmax(seg.view(12, 1024, 2), dim=-1).bool()
        flips = float((mask != mask_ref).float().mean())
        scale = float(ref.abs().max())
        err = float((out - ref).abs().max())
        print(f"SkipPropagation.generate {mode}: max|err| {err:.2e} (code scale {scale:.2f}), mask flips {flips:.2e}, "
              f"masked-in fraction {float(mask_ref.float().mean()):.2f}")
        assert out.shape == ref.shape == (2, 512, 6)
        assert 0.2 < float(mask_ref.float().mean()) < 0.8
        assert flips <= (1e-3 if mode == "x3" else 2e-2)
        assert err <= tol * max(1.0, scale) or (flips > 0 and err <= 50 * tol * max(1.0, scale))
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf


def test_detection_as_one_cuda_graph_equals_eager():
    """the whole detection pass replayed as one CUDA graph gives the eager pass's tensors bit for bit,
    for every input of the captured shape (the FPS prefix proof is a device-side flag, not a host decision)."""
    from rfdnet_b200.pipeline import GraphedDetection
    net = detection.DetectionHotPath(1, 256).eval()
    seeded_fill(net, 5)
    net = net.to(DEV)
    graphed = GraphedDetection(net)
    for seed in (11, 12, 13):
        pc = torch.from_numpy(scannet_like_batch(2, 30000, seed0=seed)).to(DEV)
        with torch.no_grad():
            ep_e, _ = net(pc)
            ep_e = {k: v.clone() for k, v in ep_e.items()}
            ep_g, _ = graphed(pc)
        for k in ("sa1_inds", "sa2_inds", "fp2_features", "vote_xyz", "aggregated_vote_inds", "objectness_scores", "center",
                  "sem_cls_scores"):
            assert torch.equal(ep_e[k], ep_g[k]), (seed, k)
    assert len(graphed.cache) == 1
    # a cloud whose first 2048 samples are NOT in FPS order for SA2 (duplicated points => ties): the flag path still agrees
    pc = torch.from_numpy(scannet_like_batch(2, 30000, seed0=14)).to(DEV)
    pc[:, 1::2] = pc[:, 0::2]
    with torch.no_grad():
        ep_e, _ = net(pc)
        ep_e = {k: v.clone() for k, v in ep_e.items()}
        ep_g, _ = graphed(pc)
    assert torch.equal(ep_e["sa2_inds"], ep_g["sa2_inds"]) and torch.equal(ep_e["objectness_scores"], ep_g["objectness_scores"])


@pytest.mark.parametrize("mode", ["x3", "fp16"])
def test_chain_ex_relu_in_group_bias_and_epilogue_pool(mode):
    """rfd_mlp_chain_ex (channel-major) and rfd_mlp_chain_rows (row-major, WideLayer): ReLU on the loaded operand, per-group
    pre-activation bias, and the max over groups of rows taken in the epilogue (values of either sign), against plain
    torch; WideLayer splits a 600-wide output into column blocks and writes at a column offset of a wider row."""
    from rfdnet_b200 import mlp
    g = torch.Generator().manual_seed(12)
    K, C, R, rows = 200, 600, 3 * 256 + 128, 128          # 7 groups of 128 rows
    G = R // rows
    W = (torch.randn(C, K, generator=g) / K ** 0.5).to(DEV)
    s = (torch.rand(C, generator=g) + 0.5).to(DEV)
    t = (torch.randn(C, generator=g) * 0.3 - 0.4).to(DEV)      # shifted down: pooled maxima of both signs
    x = torch.randn(1, K, R, generator=g).to(DEV)             # channel-major
    gb = torch.randn(1, G, C, generator=g).to(DEV)
    acc = torch.einsum("ok,kr->or", W, torch.relu(x[0])) + gb[0].t().repeat_interleave(rows, dim=1)
    ref = acc * s[:, None] + t[:, None]                       # (C, R)
    tol = (2e-5 if mode == "x3" else 5e-3) * float(ref.abs().max())
    ref_pool = ref.view(C, G, rows).amax(-1)
    assert bool((ref_pool < 0).any()) and bool((ref_pool > 0).any())
    # ---- channel-major entry, one 200-wide block of the layer
    ch = mlp.ChainMlp([(W[:200].contiguous(), s[:200].contiguous(), t[:200].contiguous(), False)], xyz=0, mode=mode)
    gbp = torch.zeros((1, G, ch.n0), device=DEV)
    gbp[:, :, :200] = gb[:, :, :200]
    pool = torch.full((1, 200, G), float("-inf"), device=DEV)
    out_cm, _ = ch.dense(x, relu_in=True, gbias=gbp, gbias_rows=rows, out_pool=pool, pool_rows=rows)
    assert float((out_cm[0] - ref[:200]).abs().max()) <= tol and float((pool[0] - ref_pool[:200]).abs().max()) <= tol
    # ---- row-major entry through WideLayer: operand = first K columns of 208-wide rows, output at column 8 of 640-wide rows
    xr = torch.full((1, R, 208), 7.0, device=DEV)
    xr[0, :, :K] = x[0].t()
    layer = mlp.WideLayer(W, s, t, False, mode)
    out = torch.full((1, R, 640), -3.0, device=DEV)
    pooled = layer(xr, out=out, out_col0=8, relu_in=True, gbias=gb, gbias_rows=rows, pool_rows=rows)
    assert float((out[0, :, 8:8 + C] - ref.t()).abs().max()) <= tol
    assert bool((out[0, :, :8] == -3.0).all()) and bool((out[0, :, 8 + C:] == -3.0).all())      # nothing else touched
    assert pooled.shape == (1, G, C) and float((pooled[0] - ref_pool.t()).abs().max()) <= tol
    # ragged row count (last tile partly empty), narrow operand and output, many tiles per CTA
    Rr = 148 * 128 * 3 + 77
    xs = torch.randn(1, Rr, 40, generator=g).to(DEV)
    Ws, ss, ts = W[:24, :40].contiguous(), s[:24].contiguous(), t[:24].contiguous()
    small = mlp.WideLayer(Ws, ss, ts, True, mode)
    outs = torch.empty((1, Rr, 24), device=DEV)
    small(xs, out=outs)
    refs = torch.relu((xs[0] @ Ws.t()) * ss[None] + ts[None])
    assert float((outs[0] - refs).abs().max()) <= (2e-5 if mode == "x3" else 5e-3) * max(1.0, float(refs.abs().max()))
    # pooled-only call (no rows written), ReLU output
    layer2 = mlp.WideLayer(W, s, t, True, mode)
    pooled2 = layer2(xr, out=None, pool_rows=rows)
    ref2 = torch.relu(torch.einsum("ok,kr->or", W, x[0]) * s[:, None] + t[:, None]).view(C, G, rows).amax(-1)
    assert float((pooled2[0] - ref2.t()).abs().max()) <= tol


def test_inplace_weight_update_in_eval_mode_is_noticed():
    """The folded / packed weight caches are keyed on the tensors' version counters: an in-place update
    while the module stays in eval() must change the fused path's output exactly like a freshly built module."""
    sa = pointnet2_modules.PointnetSAModuleVotes(npoint=128, radius=0.3, nsample=16, mlp=[5, 32, 32, 64], use_xyz=True,
                                                 normalize_xyz=True).eval()
    seeded_fill(sa, 11)
    sa = sa.to(DEV)
    xyz = torch.from_numpy(uniform_cloud(2, 1024, seed=5)).to(DEV)
    feats = torch.randn(2, 5, 1024, generator=torch.Generator().manual_seed(3)).to(DEV)
    with torch.no_grad():
        _, f0, _ = sa(xyz, feats)
        sa.mlp_module[0].weight.mul_(1.5)
        sa.mlp_module[1].running_mean.add_(0.1)
        _, f1, _ = sa(xyz, feats)
    fresh = pointnet2_modules.PointnetSAModuleVotes(npoint=128, radius=0.3, nsample=16, mlp=[5, 32, 32, 64], use_xyz=True,
                                                    normalize_xyz=True).eval().to(DEV)
    fresh.load_state_dict(sa.state_dict())
    with torch.no_grad():
        _, f2, _ = fresh(xyz, feats)
    assert not torch.equal(f0, f1) and torch.equal(f1, f2)
