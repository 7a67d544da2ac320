"""CPU restatement (torch fp32 on CPU + the C oracle for the index kernels) of the module-level hot path.
TEST INFRASTRUCTURE ONLY.  Every function takes plain tensors / state-dict style weights so that it is
independent of rfdnet_b200's nn.Modules.  Validated against the real reference Python modules by
tests/golden/make_golden.py (run in the build container, where /root/reference exists).

Citations are into /root/reference.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import torch_ext as E


def shared_mlp(x, sd, prefix, n_layers, eps=1e-5):
    """build_shared_mlp in eval mode (external/pointnet2_ops_lib/pointnet2_ops/pointnet2_modules.py:9-19):
    Conv2d 1x1 (bias=False) -> BatchNorm2d (running stats) -> ReLU, x (B,C,M,S)."""
    for i in range(n_layers):
        w = sd[f"{prefix}.{3 * i}.weight"]
        x = F.conv2d(x, w)
        bn = f"{prefix}.{3 * i + 1}"
        x = F.batch_norm(x, sd[bn + ".running_mean"], sd[bn + ".running_var"], sd[bn + ".weight"], sd[bn + ".bias"],
                         False, 0.0, eps)
        x = F.relu(x)
    return x


def query_and_group(xyz, new_xyz, features, radius, nsample, use_xyz=True, normalize_xyz=True, recip=True):
    """QueryAndGroup.forward pointnet2_utils.py:319-344.  `recip`: torch on CUDA lowers `t /= python_float` to a
    multiplication with the f32 reciprocal (the GPU reference); on CPU it is a true division."""
    idx = E.ball_query(new_xyz, xyz, radius, nsample)
    xyz_trans = xyz.transpose(1, 2).contiguous()
    grouped_xyz = E.group_points(xyz_trans, idx)
    grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
    if normalize_xyz:
        if recip:
            grouped_xyz = grouped_xyz * torch.tensor(np.float32(1.0) / np.float32(radius))
        else:
            grouped_xyz = grouped_xyz / radius
    if features is not None:
        gf = E.group_points(features.contiguous(), idx)
        new_features = torch.cat([grouped_xyz, gf], dim=1) if use_xyz else gf
    else:
        new_features = grouped_xyz
    return new_features, grouped_xyz, idx


def sa_module(xyz, features, sd, prefix, npoint, radius, nsample, n_layers=3, inds=None, recip=True):
    """PointnetSAModuleVotes.forward (pointnet2_modules.py:196-260), pooling='max', use_xyz, normalize_xyz."""
    if inds is None:
        inds = E.furthest_point_sampling(xyz.contiguous(), npoint)
    new_xyz = E.gather_points(xyz.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()
    grouped, _, _ = query_and_group(xyz.contiguous(), new_xyz, features, radius, nsample, True, True, recip)
    x = shared_mlp(grouped, sd, prefix + ".mlp_module", n_layers)
    x = F.max_pool2d(x, kernel_size=[1, x.size(3)]).squeeze(-1)
    return new_xyz, x, inds


def fp_module(unknown, known, unknow_feats, known_feats, sd, prefix, n_layers=2):
    """PointnetFPModule.forward (pointnet2_modules.py:361-405)."""
    dist2, idx = E.three_nn(unknown.contiguous(), known.contiguous())
    dist = torch.sqrt(dist2)
    dist_recip = 1.0 / (dist + 1e-8)
    norm = torch.sum(dist_recip, dim=2, keepdim=True)
    weight = dist_recip / norm
    interpolated = E.three_interpolate(known_feats.contiguous(), idx, weight.contiguous())
    x = torch.cat([interpolated, unknow_feats], dim=1) if unknow_feats is not None else interpolated
    x = shared_mlp(x.unsqueeze(-1), sd, prefix + ".mlp", n_layers)
    return x.squeeze(-1)


def _conv_bn_relu_1d(x, sd, conv, bn, eps=1e-5):
    x = F.conv1d(x, sd[conv + ".weight"], sd[conv + ".bias"])
    x = F.batch_norm(x, sd[bn + ".running_mean"], sd[bn + ".running_var"], sd[bn + ".weight"], sd[bn + ".bias"],
                     False, 0.0, eps)
    return F.relu(x)


def backbone(pointcloud, sd, prefix="backbone", recip=True):
    """Pointnet2Backbone.forward (models/iscnet/modules/pointnet2backbone.py:75-125)."""
    xyz = pointcloud[..., 0:3].contiguous()
    features = pointcloud[..., 3:].transpose(1, 2).contiguous() if pointcloud.size(-1) > 3 else None
    ep = {}
    xyz1, f1, i1 = sa_module(xyz, features, sd, prefix + ".sa1", 2048, 0.2, 64, recip=recip)
    xyz2, f2, i2 = sa_module(xyz1, f1, sd, prefix + ".sa2", 1024, 0.4, 32, recip=recip)
    xyz3, f3, _ = sa_module(xyz2, f2, sd, prefix + ".sa3", 512, 0.8, 16, recip=recip)
    xyz4, f4, _ = sa_module(xyz3, f3, sd, prefix + ".sa4", 256, 1.2, 16, recip=recip)
    f = fp_module(xyz3, xyz4, f3, f4, sd, prefix + ".fp1")
    f = fp_module(xyz2, xyz3, f2, f, sd, prefix + ".fp2")
    ep.update(sa1_inds=i1, sa1_xyz=xyz1, sa1_features=f1, sa2_inds=i2, sa2_xyz=xyz2, sa2_features=f2,
              sa3_xyz=xyz3, sa3_features=f3, sa4_xyz=xyz4, sa4_features=f4, fp2_features=f, fp2_xyz=xyz2,
              fp2_inds=i1[:, 0:xyz2.shape[1]])
    return ep


def voting(seed_xyz, seed_features, sd, prefix="voting"):
    """VotingModule.forward (vote_module.py:34-61), vote_factor 1; then network.py:323-324 L2 normalisation."""
    net = _conv_bn_relu_1d(seed_features, sd, prefix + ".conv1", prefix + ".bn1")
    net = _conv_bn_relu_1d(net, sd, prefix + ".conv2", prefix + ".bn2")
    net = F.conv1d(net, sd[prefix + ".conv3.weight"], sd[prefix + ".conv3.bias"])
    net = net.transpose(2, 1)
    vote_xyz = seed_xyz + net[:, :, 0:3]
    vote_features = (seed_features.transpose(2, 1) + net[:, :, 3:]).transpose(2, 1).contiguous()
    norm = torch.norm(vote_features, p=2, dim=1)
    vote_features = vote_features.div(norm.unsqueeze(1))
    return vote_xyz.contiguous(), vote_features


def proposal(xyz, features, sd, prefix="detection", num_proposal=256, recip=True):
    """ProposalModule.forward, sampling='vote_fps' (proposal_module.py:85-124); returns (agg_xyz, inds, net (B,69,K))."""
    xyz, features, inds = sa_module(xyz, features, sd, prefix + ".vote_aggregation", num_proposal, 0.3, 16, recip=recip)
    net = _conv_bn_relu_1d(features, sd, prefix + ".conv1", prefix + ".bn1")
    net = _conv_bn_relu_1d(net, sd, prefix + ".conv2", prefix + ".bn2")
    net = F.conv1d(net, sd[prefix + ".conv3.weight"], sd[prefix + ".conv3.bias"])
    return xyz, inds, net


def decoder(p, z, c, sd, prefix="", n_blocks=5, eps=1e-5):
    """DecoderCBatchNorm.forward in eval mode (occ_decoder.py:110-122; layers.py:98-107, 226-242)."""
    def cbn(x, name):
        gamma = F.conv1d(c.unsqueeze(2), sd[name + ".conv_gamma.weight"], sd[name + ".conv_gamma.bias"])
        beta = F.conv1d(c.unsqueeze(2), sd[name + ".conv_beta.weight"], sd[name + ".conv_beta.bias"])
        net = F.batch_norm(x, sd[name + ".bn.running_mean"], sd[name + ".bn.running_var"], None, None, False, 0.0, eps)
        return gamma * net + beta

    net = F.conv1d(p.transpose(1, 2), sd[prefix + "fc_p.weight"], sd[prefix + "fc_p.bias"])
    if (prefix + "fc_z.weight") in sd:
        net = net + F.linear(z, sd[prefix + "fc_z.weight"], sd[prefix + "fc_z.bias"]).unsqueeze(2)
    for i in range(n_blocks):
        b = f"{prefix}blocks.{i}"
        h = F.conv1d(F.relu(cbn(net, b + ".bn_0")), sd[b + ".fc_0.weight"], sd[b + ".fc_0.bias"])
        dx = F.conv1d(F.relu(cbn(h, b + ".bn_1")), sd[b + ".fc_1.weight"], sd[b + ".fc_1.bias"])
        net = net + dx
    out = F.conv1d(F.relu(cbn(net, prefix + "bn")), sd[prefix + "fc_out.weight"], sd[prefix + "fc_out.bias"])
    return out.squeeze(1)


def make_3d_grid(R=32, box_size=1.1):
    """box_size * make_3d_grid((-0.5,)*3,(0.5,)*3,(R,)*3)  (external/common.py:157-176, generator.py:92-95)."""
    lin = torch.linspace(-0.5, 0.5, R)
    px = lin.view(-1, 1, 1).expand(R, R, R).contiguous().view(-1)
    py = lin.view(1, -1, 1).expand(R, R, R).contiguous().view(-1)
    pz = lin.view(1, 1, -1).expand(R, R, R).contiguous().view(-1)
    return box_size * torch.stack([px, py, pz], dim=1)
