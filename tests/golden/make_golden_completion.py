"""Golden fixtures for the completion-side mirrors (rfdnet_b200/completion.py), produced by
the REAL reference modules (/root/reference, imported for behaviour only) on CPU:
  Encoder_Latent, ONet.compute_loss (train mode: batch-statistics CBN, KL + BCE), ResnetPointnet, PointSeg + get_loss,
  SkipPropagation.forward / .generate (STN_Group over the CPU oracle's `_ext`).
Run in the build container only: `python tests/golden/make_golden_completion.py` -> tests/golden/golden_completion.npz."""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402  (import_reference, stubs)

from rfdnet_b200.synth import scannet_like_batch, seeded_fill  # noqa: E402


class _Cfg:
    def __init__(self):
        self.config = {"data": {"z_dim": 32, "use_cls_for_completion": False, "skip_propagate": True, "c_dim": 512,
                                "threshold": 0.5, "use_color_completion": False, "no_height": False, "hidden_dim": 512}}

        class DC:
            num_class = 8
        self.dataset_config = DC()


def keys(m):
    return np.array([f"{k}:{tuple(v.shape)}" for k, v in m.state_dict().items()])


def main():
    G.import_reference()
    enc_mod = importlib.import_module("models.iscnet.modules.encoder_latent")
    layers = importlib.import_module("models.iscnet.modules.layers")
    pointseg = importlib.import_module("models.iscnet.modules.pointseg")
    onet_mod = importlib.import_module("models.iscnet.modules.occupancy_net")
    skip_mod = importlib.import_module("models.iscnet.modules.skip_propagation")
    out = {}
    g = torch.Generator().manual_seed(77)
    # ---- C1: Encoder_Latent
    enc = enc_mod.Encoder_Latent(dim=3, z_dim=32, c_dim=512)
    seeded_fill(enc, 51)
    p = torch.rand(3, 64, 3, generator=g) - 0.5
    occ = (torch.rand(3, 64, generator=g) > 0.5).float()
    c = torch.randn(3, 512, generator=g)
    with torch.no_grad():
        mean, logstd = enc(p, occ, c)
    out.update(enc_mean=mean.numpy(), enc_logstd=logstd.numpy(), keys_encoder_latent=keys(enc))
    # ---- C2: ONet.compute_loss in TRAIN mode (batch statistics), rsample seeded
    net = onet_mod.ONet(_Cfg())
    seeded_fill(net, 52)
    net.train()
    torch.manual_seed(1234)
    loss, _ = net.compute_loss(c, p, occ, None)
    loss.backward()
    out.update(onet_loss=np.float32(loss.item()),
               onet_grad_fc_p=net.decoder.fc_p.weight.grad.numpy().copy(),
               onet_grad_enc_fc_mean=net.encoder_latent.fc_mean.weight.grad.numpy().copy(),
               onet_running_mean_after=net.decoder.bn.bn.running_mean.numpy().copy(), keys_onet=keys(net))
    net.eval()
    with torch.no_grad():
        out["onet_forward_logits"] = net(p, c, None).logits.numpy()
    # ---- C3: ResnetPointnet (SkipPropagation's encoder: dim = 1 + 3 + 128, hidden 512, c_dim 512)
    rp = layers.ResnetPointnet(c_dim=512, dim=132, hidden_dim=512)
    seeded_fill(rp, 53)
    x = torch.randn(2, 50, 132, generator=g)
    with torch.no_grad():
        out["resnet_pointnet_out"] = rp(x).numpy()
    out["keys_resnet_pointnet"] = keys(rp)
    # ---- C4: PointSeg(2, 4) eval + train loss
    ps = pointseg.PointSeg(num_class=2, channel=4)
    seeded_fill(ps, 54)
    xs = torch.randn(2, 4, 128, generator=g)
    ps.eval()
    with torch.no_grad():
        lp, tf = ps(xs)
    tgt = (torch.rand(2 * 128, generator=g) > 0.5).long()
    out.update(pointseg_logp=lp.numpy(), pointseg_trans_feat=tf.numpy(),
               pointseg_loss=np.float32(pointseg.get_loss()(lp.reshape(-1, 2), tgt, tf, None).item()),
               pointseg_target=tgt.numpy(), keys_pointseg=keys(ps))
    # ---- C5: SkipPropagation.forward / generate (eval mode) on a 6000-point scene, 4 proposals
    sp = skip_mod.SkipPropagation(_Cfg())
    seeded_fill(sp, 55, scale=0.5)
    sp.eval()
    pc = torch.from_numpy(scannet_like_batch(1, 6000, seed0=66))
    box_xyz = pc[:, torch.tensor([10, 1500, 3000, 5000]), :3].contiguous() + 0.03
    orient = torch.tensor([[0.0, 0.9, -2.0, 3.1]])
    box_feat = torch.randn(1, 128, 4, generator=g)
    pil = torch.randint(0, 5, (1, 6000), generator=g).float()
    prl = torch.tensor([[1.0, 2.0, 0.0, 4.0]])
    with torch.no_grad():
        codes, mask_loss = sp(box_xyz, orient, box_feat, pc, pil, prl)
        codes_gen = sp.generate(box_xyz, orient, box_feat, pc)
    out.update(skip_codes=codes.numpy(), skip_mask_loss=np.float32(mask_loss.item()), skip_codes_generate=codes_gen.numpy(),
               skip_box_feat=box_feat.numpy(), skip_point_instance_labels=pil.numpy(), keys_skip_propagation=keys(sp))
    np.savez_compressed(os.path.join(HERE, "golden_completion.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
