"""CPU tests of the completion-side mirrors (rfdnet_b200/completion.py) against fixtures produced by the reference's own
modules (tests/golden/make_golden_completion.py): Encoder_Latent, ONet.compute_loss in train mode (batch-statistics CBN,
KL + BCE, gradients), ResnetPointnet, PointSeg + its loss, and state_dict compatibility."""
import os

import numpy as np
import pytest
import torch

from rfdnet_b200 import completion
from rfdnet_b200.synth import seeded_fill

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gc():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_completion.npz"))


def keys(m):
    return [f"{k}:{tuple(v.shape)}" for k, v in m.state_dict().items()]


def _inputs():
    g = torch.Generator().manual_seed(77)
    p = torch.rand(3, 64, 3, generator=g) - 0.5
    occ = (torch.rand(3, 64, generator=g) > 0.5).float()
    c = torch.randn(3, 512, generator=g)
    return g, p, occ, c


def test_encoder_latent_matches_reference(gc):
    enc = completion.Encoder_Latent(dim=3, z_dim=32, c_dim=512)
    assert keys(enc) == list(gc["keys_encoder_latent"])
    seeded_fill(enc, 51)
    _, p, occ, c = _inputs()
    with torch.no_grad():
        mean, logstd = enc(p, occ, c)
    assert np.allclose(mean.numpy(), gc["enc_mean"], atol=1e-5, rtol=1e-5)
    assert np.allclose(logstd.numpy(), gc["enc_logstd"], atol=1e-5, rtol=1e-5)


def test_onet_training_loss_and_gradients_match_reference(gc):
    """occupancy_net.py:59-109 in TRAIN mode: batch-statistics conditional BN, reparameterised z, KL + BCE."""
    net = completion.ONet(z_dim=32, c_dim=512)
    assert keys(net) == list(gc["keys_onet"])
    seeded_fill(net, 52)
    net.train()
    _, p, occ, c = _inputs()
    torch.manual_seed(1234)
    loss, vox = net.compute_loss(c, p, occ, None)
    loss.backward()
    assert vox is None
    assert abs(loss.item() - float(gc["onet_loss"])) <= 1e-4 * abs(float(gc["onet_loss"]))
    assert np.allclose(net.decoder.fc_p.weight.grad.numpy(), gc["onet_grad_fc_p"], atol=1e-4, rtol=1e-3)
    assert np.allclose(net.encoder_latent.fc_mean.weight.grad.numpy(), gc["onet_grad_enc_fc_mean"], atol=1e-4, rtol=1e-3)
    assert np.allclose(net.decoder.bn.bn.running_mean.numpy(), gc["onet_running_mean_after"], atol=1e-5, rtol=1e-4)
    net.eval()
    with torch.no_grad():   # CPU tensors: the reference op sequence (the sm_100a kernel needs CUDA tensors)
        logits = net(p, c, None).logits
    assert np.allclose(logits.numpy(), gc["onet_forward_logits"], atol=1e-4, rtol=1e-4)
    _, vox = net.compute_loss(c, p, occ, None, export_shape=True)
    assert vox.shape == (3, 16, 16, 16) and vox.dtype == torch.bool


def test_resnet_pointnet_matches_reference(gc):
    rp = completion.ResnetPointnet(c_dim=512, dim=132, hidden_dim=512)
    assert keys(rp) == list(gc["keys_resnet_pointnet"])
    seeded_fill(rp, 53)
    g, *_ = _inputs()
    x = torch.randn(2, 50, 132, generator=g)
    with torch.no_grad():
        out = rp(x)
    assert np.allclose(out.numpy(), gc["resnet_pointnet_out"], atol=1e-4, rtol=1e-4)


def test_pointseg_matches_reference(gc):
    ps = completion.PointSeg(num_class=2, channel=4)
    assert keys(ps) == list(gc["keys_pointseg"])
    seeded_fill(ps, 54)
    g, *_ = _inputs()
    torch.randn(2, 50, 132, generator=g)
    xs = torch.randn(2, 4, 128, generator=g)
    ps.eval()
    with torch.no_grad():
        lp, tf = ps(xs)
    assert np.allclose(lp.numpy(), gc["pointseg_logp"], atol=1e-4, rtol=1e-4)
    assert np.allclose(tf.numpy(), gc["pointseg_trans_feat"], atol=1e-4, rtol=1e-4)
    loss = completion.get_loss()(lp.reshape(-1, 2), torch.from_numpy(gc["pointseg_target"]), tf, None)
    assert abs(loss.item() - float(gc["pointseg_loss"])) <= 1e-4


def test_skip_propagation_state_dict_keys(gc):
    sp = completion.SkipPropagation(input_feature_dim=1, c_dim=512, hidden_dim=512)
    assert keys(sp) == list(gc["keys_skip_propagation"])


def test_fast_encoder_algebra_on_cpu(monkeypatch):
    """completion_fast.encode restructures PointSeg / ResnetPointnet (per-cloud biases instead of repeated global
    features, `shortcut(x) + fc_1(h)` as one layer over [x | h], epilogue max-pools, ReLU-on-load).  Here the tcgen05
    layer classes are replaced by plain torch stand-ins with the same call contract, so the ALGEBRA is checked against the
    torch mirror on the CPU in every round (the kernels themselves are checked on the GPU, tests/test_gpu_modules.py)."""
    import torch
    from rfdnet_b200 import completion, completion_fast, mlp
    from rfdnet_b200.synth import seeded_fill

    class Chain:   # contract of mlp.ChainMlp(layers, xyz=0, mode).rows(...)
        def __init__(self, layers, xyz=0, mode='x3'):
            self.layers, self.ok = layers, True
            self.K0 = layers[0][0].shape[1]
            self.out_C = layers[-1][0].shape[0]
            self.n0 = layers[0][0].shape[0]

        def rows(self, x, out=None, out_col0=0, want_out=True, relu_in=False, gbias=None, gbias_rows=0, out_pool=None,
                 pool_rows=0):
            h = x[0, :, :self.K0]
            h = torch.relu(h) if relu_in else h
            for li, (W, s, t, relu) in enumerate(self.layers):
                acc = h @ W.t()
                if li == 0 and gbias is not None:
                    acc = acc + gbias[0, :, :W.shape[0]].repeat_interleave(gbias_rows, dim=0)
                h = acc * s[None, :] + t[None, :]
                if relu:
                    h = torch.relu(h)
            if out_pool is not None:
                out_pool[0] = torch.maximum(out_pool[0], h.view(-1, pool_rows, h.shape[1]).amax(1))
            if out is None and want_out:
                return h.unsqueeze(0).clone()
            if out is not None:
                out[0, :, out_col0:out_col0 + h.shape[1]] = h
            return out

    class Wide:    # contract of mlp.WideLayer
        def __init__(self, W, scale, shift, relu, mode='x3', block=256):
            self.c = Chain([(W, scale, shift, relu)])
            self.K, self.Cout = W.shape[1], W.shape[0]

        def __call__(self, x, out=None, out_col0=0, relu_in=False, gbias=None, gbias_rows=0, pool_rows=0):
            pl = None
            if pool_rows:
                pl = torch.full((1, x.shape[1] // pool_rows, self.Cout), float("-inf"))
            self.c.rows(x, out=out, out_col0=out_col0, want_out=False, relu_in=relu_in, gbias=gbias, gbias_rows=gbias_rows,
                        out_pool=pl, pool_rows=pool_rows)
            return pl if pool_rows else out

    monkeypatch.setattr(mlp, "ChainMlp", Chain)
    monkeypatch.setattr(mlp, "WideLayer", Wide)
    sp = completion.SkipPropagation(input_feature_dim=1, c_dim=64, hidden_dim=48).eval()
    seeded_fill(sp, 5)
    g = torch.Generator().manual_seed(1)
    B, K, n = 2, 3, 128
    xyz = torch.randn(B, 3, K, n, generator=g) * 0.4
    feats = torch.cat([torch.randn(B, 1, K, n, generator=g), torch.zeros(B, 1, K, n)], dim=1)
    box = torch.randn(B, 128, K, generator=g)
    with torch.no_grad():
        # balance the segmentation head so that the mask is neither empty nor full
        pts = torch.cat([xyz, feats[:, :1]], dim=1).permute(0, 2, 3, 1).contiguous().view(B * K, n, -1)
        seg, _ = sp.point_seg(pts.transpose(1, 2).contiguous())
        lg = seg.view(-1, 2)
        d = torch.sort(lg[:, 0] - lg[:, 1]).values            # threshold half way between the two middle points:
        sp.point_seg.conv4.bias[1] += 0.5 * (d[len(d) // 2 - 1] + d[len(d) // 2])   # no point sits on the boundary
        ref, seg_pred, _ = sp._encode(xyz, feats, box)
        out, mask = completion_fast.encode(sp, xyz, feats, box, mode='x3')
    mask_ref = torch.argmax(seg_pred.view(B * K, n, 2), dim=-1).bool()
    assert 0.2 < float(mask_ref.float().mean()) < 0.8
    assert torch.equal(mask, mask_ref)
    assert out.shape == ref.shape == (B, 64, K)
    assert float((out - ref).abs().max()) <= 1e-4 * max(1.0, float(ref.abs().max()))


def test_heading_angles_match_class2angle():
    """pipeline.SceneGeneration.heading_angles == network.py:112-117 + configs/scannet_config.py:55-63 (class2angle_cuda)."""
    import math
    import torch
    from rfdnet_b200.pipeline import SceneGeneration
    g = torch.Generator().manual_seed(3)
    ep = {"heading_scores": torch.randn(2, 50, 12, generator=g), "heading_residuals_normalized": torch.randn(2, 50, 12, generator=g) * 0.4}
    got = SceneGeneration.heading_angles(ep, 12)
    cls = torch.argmax(ep["heading_scores"], -1)
    res = torch.gather(ep["heading_residuals_normalized"] * (math.pi / 12), 2, cls.unsqueeze(-1)).squeeze(2)
    angle = cls.float() * (2 * math.pi / 12.0) + res
    want = angle - 2 * math.pi * (angle > math.pi).float()
    assert torch.equal(got, want) and float(got.max()) <= math.pi
