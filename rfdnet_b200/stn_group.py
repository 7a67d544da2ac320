"""Mirrors of STN3d / STN_Group (the per-proposal grouping of SkipPropagation).

Reference: external/pointnet2_ops_lib/pointnet2_ops/pointnet2_modules.py:420-465 (STN3d), :468-537 (STN_Group);
caller models/iscnet/modules/skip_propagation.py:26-31,84-129.  Shapes in RfD-Net: ball query r = 1.0,
nsample = 1024 over the full 80k-point cloud for every kept proposal.  The reference launches its ball-query kernel
with opt_n_threads(n_proposals) threads per scene (8 threads for 10 proposals); here the same `_ext.ball_query` /
`group_points` run one warp per query, and in eval mode the three 1x1 convolutions of STN3d use the folded-BN fp32
layer kernel.  Same parameter names => reference checkpoints load.
"""
import torch
import torch.nn as nn

from . import mlp as _mlp, pointnet2_utils


class STN3d(nn.Module):
    def __init__(self, num_points=2500):
        super().__init__()
        self.num_points = num_points
        self.conv1 = nn.Conv1d(3, 64, 1)
        self.conv2 = nn.Conv1d(64, 128, 1)
        self.conv3 = nn.Conv1d(128, 256, 1)
        self.mp1 = nn.MaxPool1d(num_points)
        self.fc1 = nn.Linear(256, 128)
        self.fc2 = nn.Linear(128, 64)
        self.fc3 = nn.Linear(64, 12)
        self.relu = nn.ReLU(inplace=True)
        self.bn1 = nn.BatchNorm1d(64)
        self.bn2 = nn.BatchNorm1d(128)
        self.bn3 = nn.BatchNorm1d(256)
        self.bn4 = nn.BatchNorm1d(128)
        self.bn5 = nn.BatchNorm1d(64)
        for m in self.modules():  # weights_init (pointnet2_modules.py:407-418): zero conv / linear parameters
            if isinstance(m, (nn.Conv1d, nn.Linear)):
                nn.init.constant_(m.weight, 0.0)
                nn.init.constant_(m.bias, 0.0)

    def forward(self, grouped_xyz):
        device = grouped_xyz.device
        batch_size, _, N_proposals, _ = grouped_xyz.size()
        grouped_xyz = grouped_xyz.transpose(2, 1).contiguous().view(batch_size * N_proposals, 3, self.num_points)
        fast = not self.training and not torch.is_grad_enabled() and grouped_xyz.is_cuda
        if fast:
            x = _mlp.pointwise_layer(grouped_xyz, *_mlp.fold_conv_bn(self.conv1, self.bn1), relu=True)
            x = _mlp.pointwise_layer(x, *_mlp.fold_conv_bn(self.conv2, self.bn2), relu=True)
            x = _mlp.pointwise_layer(x, *_mlp.fold_conv_bn(self.conv3, self.bn3), relu=True)
        else:
            x = self.relu(self.bn1(self.conv1(grouped_xyz)))
            x = self.relu(self.bn2(self.conv2(x)))
            x = self.relu(self.bn3(self.conv3(x)))
        x = self.mp1(x).squeeze(2)
        x = self.relu(self.bn4(self.fc1(x)))
        x = self.relu(self.bn5(self.fc2(x)))
        x = self.fc3(x)
        iden = torch.tensor([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]]).float().view(1, 12).to(device)
        x = (x + iden).view(batch_size * N_proposals, 3, 4)
        grouped_xyz = torch.bmm(x[:, :, :3], grouped_xyz) + x[:, :, 3].unsqueeze(-1)
        grouped_xyz = grouped_xyz.view(batch_size, N_proposals, 3, -1)
        return grouped_xyz.transpose(1, 2)


class STN_Group(nn.Module):
    def __init__(self, radius=None, nsample=None, use_xyz=True, normalize_xyz=False, sample_uniformly=False,
                 ret_unique_cnt=False):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz
        self.normalize_xyz, self.ret_unique_cnt = normalize_xyz, ret_unique_cnt
        self.grouper = pointnet2_utils.QueryAndGroup(radius, nsample, use_xyz=use_xyz, ret_grouped_xyz=True,
                                                     normalize_xyz=normalize_xyz, sample_uniformly=sample_uniformly,
                                                     ret_unique_cnt=ret_unique_cnt)
        self.stn3d = STN3d(num_points=nsample)

    def forward(self, xyz, features=None, new_xyz=None, orientations=None):
        if not self.ret_unique_cnt:
            grouped_features, grouped_xyz = self.grouper(xyz, new_xyz, features)
        else:
            grouped_features, grouped_xyz, unique_cnt = self.grouper(xyz, new_xyz, features)
        rot_matrix = torch.zeros(size=[*orientations.size(), 3, 3]).to(orientations.device)
        rot_matrix[..., 0, 0] = torch.cos(orientations)
        rot_matrix[..., 0, 1] = torch.sin(orientations)
        rot_matrix[..., 1, 1] = torch.cos(orientations)
        rot_matrix[..., 1, 0] = -torch.sin(orientations)
        rot_matrix[..., 2, 2] = 1.
        batch_size, N_proposals = orientations.size()
        grouped_xyz = torch.bmm(rot_matrix.view(batch_size * N_proposals, 3, 3),
                                grouped_xyz.transpose(1, 2).contiguous().view(batch_size * N_proposals, 3, -1))
        grouped_xyz = grouped_xyz.view(batch_size, N_proposals, 3, -1).transpose(1, 2).contiguous()
        grouped_xyz = self.stn3d(grouped_xyz)
        if not self.ret_unique_cnt:
            return grouped_xyz, grouped_features
        return grouped_xyz, grouped_features, unique_cnt
