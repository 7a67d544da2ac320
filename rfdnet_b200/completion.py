"""Host-side mirrors of the shape-completion half of ISCNet -- the training-side ONet and
SkipPropagation's dense networks.  The reference's host side is Python/PyTorch, so is this; parameter names and shapes
equal the reference's (state_dict compatible), the dense layers run on PyTorch's library GEMMs (training needs
autograd), the point-cloud operators underneath (STN_Group: ball query r = 1.0 / nsample = 1024, grouping and their
scatter-add gradients) on librfdnet_b200.

  Encoder_Latent   models/iscnet/modules/encoder_latent.py:12-72
  ONet             models/iscnet/modules/occupancy_net.py:11-189 (compute_loss :59-109, decode :147-156, infer_z :158-175)
  ResnetBlockFC    models/iscnet/modules/layers.py:9-48
  ResnetPointnet   models/iscnet/modules/layers.py:340-392
  STN3d / STNkd / PointNetEncoder / PointSeg / get_loss   models/iscnet/modules/pointseg.py:7-182
  SkipPropagation  models/iscnet/modules/skip_propagation.py:13-129
"""
import torch
import torch.distributions as dist
import torch.nn as nn
import torch.nn.functional as F

from . import onet
from .stn_group import STN_Group


def _maxpool(x, dim=-1, keepdim=False):
    return x.max(dim=dim, keepdim=keepdim)[0]


def _append_pooled(net):
    """(B,T,F) -> (B,T,2F): every point gets the max over the T points appended (the PointNet 'global' half)"""
    return torch.cat([net, _maxpool(net, dim=1, keepdim=True).expand_as(net)], dim=2)


class Encoder_Latent(nn.Module):
    """q(z | p, occ, c): encoder_latent.py:12-72 (leaky=False)."""

    def __init__(self, z_dim=128, c_dim=128, dim=3):
        super().__init__()
        self.z_dim, self.c_dim = z_dim, c_dim
        self.fc_pos = nn.Linear(dim, 128)
        if c_dim != 0:
            self.fc_c = nn.Linear(c_dim, 128)
        self.fc_0 = nn.Linear(1, 128)
        self.fc_1 = nn.Linear(128, 128)
        self.fc_2 = nn.Linear(256, 128)
        self.fc_3 = nn.Linear(256, 128)
        self.fc_mean = nn.Linear(128, z_dim)
        self.fc_logstd = nn.Linear(128, z_dim)

    def forward(self, p, x, c=None, **kwargs):
        net = self.fc_0(x.unsqueeze(-1)) + self.fc_pos(p)
        if self.c_dim != 0:
            net = net + self.fc_c(c).unsqueeze(1)
        net = self.fc_1(F.relu(net))
        net = self.fc_2(F.relu(_append_pooled(net)))
        net = self.fc_3(F.relu(_append_pooled(net)))
        net = _maxpool(net, dim=1)
        return self.fc_mean(net), self.fc_logstd(net)


class ONet(nn.Module):
    """occupancy_net.py:11-189 with the cfg lookups replaced by keyword arguments (ISCNet.yaml: z_dim 32, c_dim 512,
    use_cls_for_completion False, threshold 0.5).  `decoder` is the same DecoderCBatchNorm mirror the inference path uses:
    in train mode / under autograd it runs the reference's op sequence (batch-statistics CBN), in eval mode the tcgen05
    kernel."""

    def __init__(self, z_dim=32, c_dim=512, num_class=8, use_cls_for_completion=False, threshold=0.5, precision='fp16'):
        super().__init__()
        self.z_dim, self.threshold = z_dim, threshold
        self.use_cls_for_completion = use_cls_for_completion
        c_dim = int(use_cls_for_completion) * num_class + c_dim
        self.encoder_latent = Encoder_Latent(dim=3, z_dim=z_dim, c_dim=c_dim) if z_dim != 0 else None
        self.decoder = onet.DecoderCBatchNorm(dim=3, z_dim=z_dim, c_dim=c_dim, precision=precision)

    def get_prior_z(self, z_dim, device):
        return dist.Normal(torch.zeros(z_dim, device=device), torch.ones(z_dim, device=device))

    def get_z_from_prior(self, size=torch.Size([]), device='cuda', sample=False):
        p0_z = self.get_prior_z(self.z_dim, device)
        if sample:
            return p0_z.sample(size)
        return p0_z.mean.expand(*size, *p0_z.mean.size())

    def infer_z(self, p, occ, c, device, **kwargs):
        if self.encoder_latent is not None:
            mean_z, logstd_z = self.encoder_latent(p, occ, c, **kwargs)
        else:
            mean_z = torch.empty(p.size(0), 0, device=device)
            logstd_z = torch.empty(p.size(0), 0, device=device)
        return dist.Normal(mean_z, torch.exp(logstd_z))

    def decode(self, input_points_for_completion, z, features, **kwargs):
        return dist.Bernoulli(logits=self.decoder(input_points_for_completion, z, features, **kwargs))

    def _with_cls(self, features, cls_codes):
        if self.use_cls_for_completion:
            features = torch.cat([features, cls_codes.to(features.device).float()], dim=-1)
        return features

    def compute_loss(self, input_features_for_completion, input_points_for_completion,
                     input_points_occ_for_completion, cls_codes_for_completion=None, export_shape=False):
        """KL(q(z|.) || N(0,1)) + BCE-with-logits summed over the points, averaged over the boxes (:59-109)."""
        feats = self._with_cls(input_features_for_completion, cls_codes_for_completion)
        device, nbox = feats.device, feats.size(0)
        if self.z_dim > 0:
            q_z = self.infer_z(input_points_for_completion, input_points_occ_for_completion, feats, device)
            z = q_z.rsample()
            loss = dist.kl_divergence(q_z, self.get_prior_z(self.z_dim, device)).sum(dim=-1).mean()
        else:
            z = torch.empty(size=(nbox, 0), device=device)
            loss = 0.
        logits = self.decode(input_points_for_completion, z, feats).logits
        bce = F.binary_cross_entropy_with_logits(logits, input_points_occ_for_completion, reduction='none')
        loss = loss + bce.sum(-1).mean()
        voxels_out = None
        if export_shape:
            shape = (16, 16, 16)
            p = onet.make_3d_grid_cpu([-0.5 + 1 / 32] * 3, [0.5 - 1 / 32] * 3, shape).to(device)
            p = p.expand(nbox, *p.size())
            z0 = self.get_z_from_prior((nbox,), device, sample=False)
            voxels_out = self.decode(p, z0, feats).probs.view(nbox, *shape) >= self.threshold
        return loss, voxels_out

    def forward(self, input_points_for_completion, input_features_for_completion, cls_codes_for_completion=None,
                sample=False, **kwargs):
        feats = self._with_cls(input_features_for_completion, cls_codes_for_completion)
        z = self.get_z_from_prior((input_points_for_completion.size(0),), feats.device, sample=sample)
        return self.decode(input_points_for_completion, z, feats, **kwargs)


class ResnetBlockFC(nn.Module):
    """layers.py:9-48: x_s + fc_1(relu(fc_0(relu(x)))), fc_1 zero-initialised.  The reference's activation is
    nn.ReLU(inplace=True) (:30), so by the time its shortcut reads `x` (:41-44) the tensor already holds relu(x): the
    residual branch is shortcut(relu(x)) (or relu(x) itself when the widths agree).  Reproduced here without the aliasing."""

    def __init__(self, size_in, size_out=None, size_h=None):
        super().__init__()
        size_out = size_in if size_out is None else size_out
        size_h = min(size_in, size_out) if size_h is None else size_h
        self.size_in, self.size_h, self.size_out = size_in, size_h, size_out
        self.fc_0 = nn.Linear(size_in, size_h)
        self.fc_1 = nn.Linear(size_h, size_out)
        self.actvn = nn.ReLU()
        self.shortcut = None if size_in == size_out else nn.Linear(size_in, size_out, bias=False)
        nn.init.zeros_(self.fc_1.weight)

    def forward(self, x):
        x = self.actvn(x)
        dx = self.fc_1(self.actvn(self.fc_0(x)))
        return (x if self.shortcut is None else self.shortcut(x)) + dx


class ResnetPointnet(nn.Module):
    """layers.py:340-392: five residual blocks with the max-pooled global feature appended between them."""

    def __init__(self, c_dim=128, dim=3, hidden_dim=128):
        super().__init__()
        self.c_dim = c_dim
        self.fc_pos = nn.Linear(dim, 2 * hidden_dim)
        for i in range(5):
            setattr(self, f"block_{i}", ResnetBlockFC(2 * hidden_dim, hidden_dim))
        self.fc_c = nn.Linear(hidden_dim, c_dim)
        self.actvn = nn.ReLU()

    def forward(self, p):
        net = self.block_0(self.fc_pos(p))
        for i in range(1, 5):
            net = getattr(self, f"block_{i}")(_append_pooled(net))
        return self.fc_c(self.actvn(_maxpool(net, dim=1)))


class _TNet(nn.Module):
    """pointseg.py:7-85 (STN3d / STNkd): per-cloud k x k alignment matrix, identity added to the regressed one."""

    def __init__(self, cin, k):
        super().__init__()
        self.k = k
        self.conv1, self.conv2, self.conv3 = nn.Conv1d(cin, 64, 1), nn.Conv1d(64, 128, 1), nn.Conv1d(128, 1024, 1)
        self.fc1, self.fc2, self.fc3 = nn.Linear(1024, 512), nn.Linear(512, 256), nn.Linear(256, k * k)
        self.relu = nn.ReLU()
        self.bn1, self.bn2, self.bn3 = nn.BatchNorm1d(64), nn.BatchNorm1d(128), nn.BatchNorm1d(1024)
        self.bn4, self.bn5 = nn.BatchNorm1d(512), nn.BatchNorm1d(256)

    def forward(self, x):
        for conv, bn in ((self.conv1, self.bn1), (self.conv2, self.bn2), (self.conv3, self.bn3)):
            x = F.relu(bn(conv(x)))
        x = x.max(dim=2)[0]
        x = F.relu(self.bn4(self.fc1(x)))
        x = F.relu(self.bn5(self.fc2(x)))
        x = self.fc3(x) + torch.eye(self.k, device=x.device, dtype=x.dtype).reshape(1, -1)
        return x.view(-1, self.k, self.k)


class STN3d(_TNet):
    def __init__(self, channel):
        super().__init__(channel, 3)


class STNkd(_TNet):
    def __init__(self, k=64):
        super().__init__(k, k)


class PointNetEncoder(nn.Module):
    """pointseg.py:88-133"""

    def __init__(self, global_feat=True, feature_transform=False, channel=3):
        super().__init__()
        self.stn = STN3d(channel)
        self.conv1, self.conv2, self.conv3 = nn.Conv1d(channel, 64, 1), nn.Conv1d(64, 128, 1), nn.Conv1d(128, 1024, 1)
        self.bn1, self.bn2, self.bn3 = nn.BatchNorm1d(64), nn.BatchNorm1d(128), nn.BatchNorm1d(1024)
        self.global_feat, self.feature_transform = global_feat, feature_transform
        if feature_transform:
            self.fstn = STNkd(k=64)

    def forward(self, x):
        B, D, N = x.size()
        trans = self.stn(x)
        x = x.transpose(2, 1)
        xyz = torch.bmm(x[..., :3], trans)             # only the coordinates are rotated
        x = (torch.cat([xyz, x[..., 3:]], dim=2) if D > 3 else xyz).transpose(2, 1)
        x = F.relu(self.bn1(self.conv1(x)))
        trans_feat = None
        if self.feature_transform:
            trans_feat = self.fstn(x)
            x = torch.bmm(x.transpose(2, 1), trans_feat).transpose(2, 1)
        pointfeat = x
        x = F.relu(self.bn2(self.conv2(x)))
        x = self.bn3(self.conv3(x))
        x = x.max(dim=2, keepdim=True)[0].view(-1, 1024)
        if self.global_feat:
            return x, trans, trans_feat
        return torch.cat([x.view(-1, 1024, 1).repeat(1, 1, N), pointfeat], 1), trans, trans_feat


def feature_transform_reguliarzer(trans):
    """pointseg.py:135-142 (the reference multiplies trans with (trans^T - I); kept as is)"""
    eye = torch.eye(trans.size(1), device=trans.device)[None]
    return torch.mean(torch.norm(torch.bmm(trans, trans.transpose(2, 1) - eye), dim=(1, 2)))


class PointSeg(nn.Module):
    """pointseg.py:144-168: per-point 2-class log-probabilities"""

    def __init__(self, num_class, channel):
        super().__init__()
        self.k = num_class
        self.feat = PointNetEncoder(global_feat=False, feature_transform=True, channel=channel)
        self.conv1, self.conv2 = nn.Conv1d(1088, 512, 1), nn.Conv1d(512, 256, 1)
        self.conv3, self.conv4 = nn.Conv1d(256, 128, 1), nn.Conv1d(128, self.k, 1)
        self.bn1, self.bn2, self.bn3 = nn.BatchNorm1d(512), nn.BatchNorm1d(256), nn.BatchNorm1d(128)

    def forward(self, x):
        B, _, n_pts = x.size()
        x, _, trans_feat = self.feat(x)
        for conv, bn in ((self.conv1, self.bn1), (self.conv2, self.bn2), (self.conv3, self.bn3)):
            x = F.relu(bn(conv(x)))
        x = self.conv4(x).transpose(2, 1).contiguous()
        return F.log_softmax(x.view(-1, self.k), dim=-1).view(B, n_pts, self.k), trans_feat


class get_loss(nn.Module):
    """pointseg.py:171-182"""

    def __init__(self, mat_diff_loss_scale=0.001):
        super().__init__()
        self.mat_diff_loss_scale = mat_diff_loss_scale

    def forward(self, pred, target, trans_feat, weight):
        return F.nll_loss(pred, target, weight=weight) + feature_transform_reguliarzer(trans_feat) * self.mat_diff_loss_scale


class SkipPropagation(nn.Module):
    """skip_propagation.py:13-129: for every kept proposal, group <= 1024 scene points within 1 m of the box centre in
    the box frame (STN_Group on the sm_100a kernels), segment them (PointSeg), mask, and encode them together with the
    proposal feature into the 512-d shape code the ONet decoder is conditioned on (ResnetPointnet)."""

    def __init__(self, input_feature_dim=1, c_dim=512, hidden_dim=512, proposal_dim=128):
        super().__init__()
        self.input_feature_dim = input_feature_dim
        self.stn = STN_Group(radius=1., nsample=1024, use_xyz=False, normalize_xyz=True)
        self.encoder = ResnetPointnet(c_dim=c_dim, dim=input_feature_dim + 3 + proposal_dim, hidden_dim=hidden_dim)
        self.point_seg = PointSeg(num_class=2, channel=input_feature_dim + 3)
        self.mask_loss_func = get_loss()

    def _break_up_pc(self, pc):
        xyz = pc[..., 0:3].contiguous()
        feats = pc[..., 3:3 + self.input_feature_dim].transpose(1, 2).contiguous() if pc.size(-1) > 3 else None
        return xyz, feats

    def _encode(self, xyz, feats, box_feature):
        """shared tail of forward / generate: PointSeg mask -> masked [xyz, feature, proposal feature] -> encoder"""
        B, _, K, n = feats.size()
        pts = torch.cat([xyz, feats[:, :1]], dim=1).permute(0, 2, 3, 1).contiguous().view(B * K, n, -1)
        seg_pred, trans_feat = self.point_seg(pts.transpose(1, 2).contiguous())
        seg_pred = seg_pred.contiguous().view(B * K * n, 2)
        box = box_feature.transpose(1, 2).contiguous().view(B * K, 1, -1).expand(-1, n, -1)
        x = torch.cat([pts, box], dim=2)
        mask = torch.argmax(seg_pred, dim=1).view(B * K, n, 1)
        x = x * mask.float()
        return self.encoder(x).view(B, K, -1).transpose(1, 2), seg_pred, trans_feat

    def generate(self, box_xyz, box_orientations, box_feature, input_point_cloud):
        """eval + no_grad + CUDA: STN_Group on its four fused launches, PointSeg and ResnetPointnet on the tcgen05 chain
        kernel (completion_fast.encode; `self.fast_precision`: 'x3' fp32-grade default, 'fp16', or None = torch layers)."""
        xyz, feats = self._break_up_pc(input_point_cloud)
        feats = torch.cat([feats, torch.zeros_like(feats)], dim=1)  # instance labels are not used in generation
        xyz, feats = self.stn(xyz, feats, box_xyz, box_orientations)
        mode = getattr(self, "fast_precision", "x3")
        if (mode is not None and not self.training and not torch.is_grad_enabled() and xyz.is_cuda
                and xyz.shape[-1] % 128 == 0 and self.encoder.block_0.size_h <= 512):
            from . import completion_fast
            return completion_fast.encode(self, xyz, feats, box_feature, mode)[0]
        return self._encode(xyz, feats, box_feature)[0]

    def forward(self, box_xyz, box_orientations, box_feature, input_point_cloud, point_instance_labels,
                proposal_instance_labels):
        xyz, feats = self._break_up_pc(input_point_cloud)
        feats = torch.cat([feats, point_instance_labels.unsqueeze(1).to(feats.dtype)], dim=1)
        xyz, feats = self.stn(xyz, feats, box_xyz, box_orientations)
        n = feats.size(3)
        target = (feats[:, 1] == proposal_instance_labels.unsqueeze(-1).to(feats.dtype)).reshape(-1)
        codes, seg_pred, trans_feat = self._encode(xyz, feats, box_feature)
        return codes, self.mask_loss_func(seg_pred, target.long(), trans_feat, weight=None)
