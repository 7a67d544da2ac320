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
