"""Host-side mirrors of the detection half of ISCNet:
  Pointnet2Backbone  models/iscnet/modules/pointnet2backbone.py:11-125   (4 SA + 2 FP)
  VotingModule       models/iscnet/modules/vote_module.py:12-61
  ProposalModule     models/iscnet/modules/proposal_module.py:43-124 (+ decode_scores :13-39)
  vote feature L2 normalisation  models/iscnet/modules/network.py:323-324
Same attribute names => same state_dict keys as the reference modules.  The reference's `cfg`
object is replaced by explicit keyword arguments carrying the same values
(configs/config_files/ISCNet.yaml, configs/scannet_config.py:13-15).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import mlp as _mlp, pointnet2_utils
from .pointnet2_modules import PointnetFPModule, PointnetSAModuleVotes


class Pointnet2Backbone(nn.Module):
    def __init__(self, input_feature_dim=1):
        super().__init__()
        self.input_feature_dim = input_feature_dim
        self.sa1 = PointnetSAModuleVotes(npoint=2048, radius=0.2, nsample=64,
                                         mlp=[self.input_feature_dim, 64, 64, 128], use_xyz=True, normalize_xyz=True)
        self.sa2 = PointnetSAModuleVotes(npoint=1024, radius=0.4, nsample=32, mlp=[128, 128, 128, 256],
                                         use_xyz=True, normalize_xyz=True)
        self.sa3 = PointnetSAModuleVotes(npoint=512, radius=0.8, nsample=16, mlp=[256, 128, 128, 256],
                                         use_xyz=True, normalize_xyz=True)
        self.sa4 = PointnetSAModuleVotes(npoint=256, radius=1.2, nsample=16, mlp=[256, 128, 128, 256],
                                         use_xyz=True, normalize_xyz=True)
        self.fp1 = PointnetFPModule(mlp=[256 + 256, 256, 256])
        self.fp2 = PointnetFPModule(mlp=[256 + 256, 256, 256])

    def _forward_fused(self, xyz, features, end_points):
        """Inference schedule: every SA layer hands its features to the next one in point-major layout (the layout
        the gather of the tensor-core chain kernel reads), next to the channel-major tensors the reference exposes."""
        i1, x1 = pointnet2_utils.fps_with_xyz(xyz.contiguous(), self.sa1.npoint)
        _, f1, _, p1 = self.sa1._forward_fused(xyz, features, i1, x1, want_pm=True)
        # SA2-4 sample an FPS-ordered cloud: the identity result is proved per scene instead of re-running the sampler
        i2, x2 = pointnet2_utils.fps_with_xyz(x1, self.sa2.npoint, try_prefix=True)
        i3, x3 = pointnet2_utils.fps_with_xyz(x2, self.sa3.npoint, try_prefix=True)
        i4, x4 = pointnet2_utils.fps_with_xyz(x3, self.sa4.npoint, try_prefix=True)
        _, f2, _, p2 = self.sa2._forward_fused(x1, f1, i2, x2, features_pm=p1, want_pm=True)
        _, f3, _, p3 = self.sa3._forward_fused(x2, f2, i3, x3, features_pm=p2, want_pm=True)
        _, f4, _, _ = self.sa4._forward_fused(x3, f3, i4, x4, features_pm=p3)
        f = self.fp1(x3, x4, f3, f4)
        f = self.fp2(x2, x3, f2, f)
        end_points.update(sa1_inds=i1, sa1_xyz=x1, sa1_features=f1, sa2_inds=i2, sa2_xyz=x2, sa2_features=f2,
                          sa3_xyz=x3, sa3_features=f3, sa4_xyz=x4, sa4_features=f4, fp2_features=f, fp2_xyz=x2,
                          fp2_inds=i1[:, 0:x2.shape[1]])
        return end_points

    def _break_up_pc(self, pc):
        xyz = pc[..., 0:3].contiguous()
        features = (pc[..., 3:3 + self.input_feature_dim].transpose(1, 2).contiguous()
                    if pc.size(-1) > 3 else None)
        return xyz, features

    def forward(self, pointcloud, end_points=None):
        if not end_points:
            end_points = {}
        xyz, features = self._break_up_pc(pointcloud)
        if (not self.training and not torch.is_grad_enabled() and xyz.is_cuda
                and all(getattr(self, n)._fast_ok(xyz, None) for n in ("sa1", "sa2", "sa3", "sa4"))):
            return self._forward_fused(xyz, features, end_points)
        xyz, features, fps_inds = self.sa1(xyz, features)
        end_points['sa1_inds'] = fps_inds
        end_points['sa1_xyz'] = xyz
        end_points['sa1_features'] = features
        xyz, features, fps_inds = self.sa2(xyz, features)
        end_points['sa2_inds'] = fps_inds
        end_points['sa2_xyz'] = xyz
        end_points['sa2_features'] = features
        xyz, features, fps_inds = self.sa3(xyz, features)
        end_points['sa3_xyz'] = xyz
        end_points['sa3_features'] = features
        xyz, features, fps_inds = self.sa4(xyz, features)
        end_points['sa4_xyz'] = xyz
        end_points['sa4_features'] = features
        features = self.fp1(end_points['sa3_xyz'], end_points['sa4_xyz'], end_points['sa3_features'],
                            end_points['sa4_features'])
        features = self.fp2(end_points['sa2_xyz'], end_points['sa3_xyz'], end_points['sa2_features'], features)
        end_points['fp2_features'] = features
        end_points['fp2_xyz'] = end_points['sa2_xyz']
        num_seed = end_points['fp2_xyz'].shape[1]
        end_points['fp2_inds'] = end_points['sa1_inds'][:, 0:num_seed]
        return end_points


class _HeadFoldCache:
    """folded (W, scale, shift) of conv1/bn1, conv2/bn2, conv3 -- recomputed only after the weights change"""
    precision = 'x3'  # 'x3' | 'fp16' | 'bf16': one tcgen05 chain kernel for the three convs; 'cuda': fp32 layer kernel

    def _heads(self):
        ver = tuple(_mlp.state_version(m) for m in (self.conv1, self.bn1, self.conv2, self.bn2, self.conv3))
        if getattr(self, "_hf", None) is None or getattr(self, "_hf_ver", None) != ver:   # also after in-place updates
            self._drop_heads()
            self._hf = (_mlp.fold_conv_bn(self.conv1, self.bn1), _mlp.fold_conv_bn(self.conv2, self.bn2),
                        _mlp.fold_conv_bn(self.conv3, None))
            self._hf_ver = ver
        return self._hf

    def _run_heads(self, x):
        """conv1+bn1+relu -> conv2+bn2+relu -> conv3 on x (B,C,L) f32 contiguous"""
        h1, h2, h3 = self._heads()
        if self.precision != 'cuda':
            tc = getattr(self, "_tc", None)
            if tc is None or tc[0] != self.precision:
                tc = self._tc = (self.precision, _mlp.ChainMlp([(*h1, True), (*h2, True), (*h3, False)], xyz=0,
                                                               mode=self.precision))
            if tc[1].ok:
                return tc[1].dense(x)[0]
        net = _mlp.pointwise_layer(x, *h1, relu=True)
        net = _mlp.pointwise_layer(net, *h2, relu=True)
        return _mlp.pointwise_layer(net, *h3, relu=False)

    def _drop_heads(self):
        self._hf = None
        self._tc = None

    def train(self, mode=True):
        self._drop_heads()
        return super().train(mode)

    def _load_from_state_dict(self, *a, **k):
        self._drop_heads()
        return super()._load_from_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._drop_heads()
        return super()._apply(fn, *a, **k)


class VotingModule(_HeadFoldCache, nn.Module):
    def __init__(self, vote_factor=1, seed_feature_dim=256):
        super().__init__()
        self.vote_factor = vote_factor
        self.in_dim = seed_feature_dim
        self.out_dim = self.in_dim
        self.conv1 = nn.Conv1d(self.in_dim, self.in_dim, 1)
        self.conv2 = nn.Conv1d(self.in_dim, self.in_dim, 1)
        self.conv3 = nn.Conv1d(self.in_dim, (3 + self.out_dim) * self.vote_factor, 1)
        self.bn1 = nn.BatchNorm1d(self.in_dim)
        self.bn2 = nn.BatchNorm1d(self.in_dim)

    def forward(self, seed_xyz, seed_features):
        batch_size, num_seed = seed_xyz.shape[0], seed_xyz.shape[1]
        num_vote = num_seed * self.vote_factor
        fast = (not self.training and not torch.is_grad_enabled() and seed_features.is_cuda
                and seed_features.dtype == torch.float32 and self.conv1.weight.device == seed_features.device)
        if fast:
            net = self._run_heads(seed_features.contiguous())
        else:
            net = F.relu(self.bn1(self.conv1(seed_features)))
            net = F.relu(self.bn2(self.conv2(net)))
            net = self.conv3(net)
        net = net.transpose(2, 1).view(batch_size, num_seed, self.vote_factor, 3 + self.out_dim)
        offset = net[:, :, :, 0:3]
        vote_xyz = (seed_xyz.unsqueeze(2) + offset).contiguous().view(batch_size, num_vote, 3)
        residual_features = net[:, :, :, 3:]
        vote_features = seed_features.transpose(2, 1).unsqueeze(2) + residual_features
        vote_features = vote_features.contiguous().view(batch_size, num_vote, self.out_dim)
        vote_features = vote_features.transpose(2, 1).contiguous()
        return vote_xyz, vote_features


def decode_scores(net, end_points, num_heading_bin, num_size_cluster):
    """proposal_module.py:13-39"""
    net_transposed = net.transpose(2, 1)
    batch_size, num_proposal = net_transposed.shape[0], net_transposed.shape[1]
    end_points['objectness_scores'] = net_transposed[:, :, 0:2]
    base_xyz = end_points['aggregated_vote_xyz']
    end_points['center'] = base_xyz + net_transposed[:, :, 2:5]
    end_points['heading_scores'] = net_transposed[:, :, 5:5 + num_heading_bin]
    end_points['heading_residuals_normalized'] = net_transposed[:, :, 5 + num_heading_bin:5 + num_heading_bin * 2]
    end_points['size_scores'] = net_transposed[:, :, 5 + num_heading_bin * 2:5 + num_heading_bin * 2 + num_size_cluster]
    end_points['size_residuals_normalized'] = net_transposed[
        :, :, 5 + num_heading_bin * 2 + num_size_cluster:5 + num_heading_bin * 2 + num_size_cluster * 4].reshape(
        [batch_size, num_proposal, num_size_cluster, 3])
    end_points['sem_cls_scores'] = net_transposed[:, :, 5 + num_heading_bin * 2 + num_size_cluster * 4:]
    return end_points


class ProposalModule(_HeadFoldCache, nn.Module):
    def __init__(self, num_class=8, num_heading_bin=12, num_size_cluster=8, num_proposal=256,
                 sampling='vote_fps', seed_feat_dim=256):
        super().__init__()
        self.num_class, self.num_heading_bin, self.num_size_cluster = num_class, num_heading_bin, num_size_cluster
        self.num_proposal, self.sampling, self.seed_feat_dim = num_proposal, sampling, seed_feat_dim
        self.vote_aggregation = PointnetSAModuleVotes(npoint=self.num_proposal, radius=0.3, nsample=16,
                                                      mlp=[self.seed_feat_dim, 128, 128, 128], use_xyz=True,
                                                      normalize_xyz=True)
        self.conv1 = nn.Conv1d(128, 128, 1)
        self.conv2 = nn.Conv1d(128, 128, 1)
        self.conv3 = nn.Conv1d(128, 2 + 3 + num_heading_bin * 2 + num_size_cluster * 4 + self.num_class, 1)
        self.bn1 = nn.BatchNorm1d(128)
        self.bn2 = nn.BatchNorm1d(128)

    def forward(self, xyz, features, end_points, export_proposal_feature=False):
        if self.sampling == 'vote_fps':
            xyz, features, fps_inds = self.vote_aggregation(xyz, features)
            sample_inds = fps_inds
        elif self.sampling == 'seed_fps':
            sample_inds = pointnet2_utils.furthest_point_sample(end_points['seed_xyz'], self.num_proposal)
            xyz, features, _ = self.vote_aggregation(xyz, features, sample_inds)
        elif self.sampling == 'random':
            num_seed, batch_size = end_points['seed_xyz'].shape[1], end_points['seed_xyz'].shape[0]
            sample_inds = torch.randint(0, num_seed, (batch_size, self.num_proposal), dtype=torch.int,
                                        device=xyz.device)
            xyz, features, _ = self.vote_aggregation(xyz, features, sample_inds)
        else:
            raise ValueError('Unknown sampling strategy: %s' % self.sampling)
        end_points['aggregated_vote_xyz'] = xyz
        end_points['aggregated_vote_inds'] = sample_inds
        fast = (not self.training and not torch.is_grad_enabled() and features.is_cuda
                and features.dtype == torch.float32 and self.conv1.weight.device == features.device)
        if fast:
            net = self._run_heads(features.contiguous())
        else:
            net = F.relu(self.bn1(self.conv1(features)))
            net = F.relu(self.bn2(self.conv2(net)))
            net = self.conv3(net)
        end_points = decode_scores(net, end_points, self.num_heading_bin, self.num_size_cluster)
        return (end_points, features) if export_proposal_feature else (end_points, None)


class DetectionHotPath(nn.Module):
    """backbone -> voting -> L2 norm -> proposal: ISCNet.forward lines network.py:313-330."""

    def __init__(self, input_feature_dim=1, num_proposal=256, **proposal_kw):
        super().__init__()
        self.backbone = Pointnet2Backbone(input_feature_dim)
        self.voting = VotingModule(1, 256)
        self.detection = ProposalModule(num_proposal=num_proposal, **proposal_kw)

    def forward(self, point_clouds, export_proposal_feature=False):
        end_points = self.backbone(point_clouds, {})
        xyz, features = end_points['fp2_xyz'], end_points['fp2_features']
        end_points['seed_inds'] = end_points['fp2_inds']
        end_points['seed_xyz'] = xyz
        end_points['seed_features'] = features
        xyz, features = self.voting(xyz, features)
        features_norm = torch.norm(features, p=2, dim=1)
        features = features.div(features_norm.unsqueeze(1))
        end_points['vote_xyz'] = xyz
        end_points['vote_features'] = features
        end_points, proposal_features = self.detection(xyz, features, end_points, export_proposal_feature)
        return end_points, proposal_features
