"""Host-side mirror of the reference SA / FP modules
(external/pointnet2_ops_lib/pointnet2_ops/pointnet2_modules.py: build_shared_mlp :9-19,
PointnetSAModuleVotes :149-260, PointnetFPModule :345-405).

Same constructor arguments, forward signatures, return values and state_dict keys
(`mlp_module.{0,3,6}.weight`, `mlp_module.{1,4,7}.*`, `mlp.*`), so reference checkpoints load.
In eval mode without autograd the forward runs entirely on the sm_100a kernels
(FPS -> ball query -> gather + folded-BN MLP + max in one tcgen05 kernel); with autograd enabled (training,
config 5) the reference's op sequence runs on the drop-in `_ext` kernels + torch autograd.
"""
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, mlp as _mlp, pointnet2_utils


def build_shared_mlp(mlp_spec: List[int], bn: bool = True):
    layers = []
    for i in range(1, len(mlp_spec)):
        layers.append(nn.Conv2d(mlp_spec[i - 1], mlp_spec[i], kernel_size=1, bias=not bn))
        if bn:
            layers.append(nn.BatchNorm2d(mlp_spec[i]))
        layers.append(nn.ReLU(True))
    return nn.Sequential(*layers)


class _FoldCache:
    """Lazily folded (W, scale, shift, relu) per layer and the packed tensor-core image built from them; rebuilt whenever
    a parameter or buffer of the MLP changes (mode switch, load_state_dict, .to(), or an in-place update: the cache is
    keyed on the tensors' version counters)."""

    def _folded(self, seq):
        ver = _mlp.state_version(seq)
        if getattr(self, "_fold", None) is None or getattr(self, "_fold_ver", None) != ver:
            self._drop()
            self._fold = _mlp.fold_sequential(seq)
            self._fold_ver = ver
        return self._fold

    def _drop(self):
        self._fold = None
        self._tc = None

    def train(self, mode=True):
        self._drop()
        return super().train(mode)

    def _load_from_state_dict(self, *a, **k):
        self._drop()
        return super()._load_from_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._drop()
        return super()._apply(fn, *a, **k)


class PointnetSAModuleVotes(_FoldCache, nn.Module):
    """pointnet2_modules.py:149-260 (pooling 'max' | 'avg' | 'rbf').

    precision (inference path): 'x3' (default) / 'fp16' / 'bf16' run FPS -> ball query -> ONE tcgen05 kernel that
    gathers, centres, applies the three folded Conv+BN+ReLU layers and max-pools (csrc/mlp_chain_tc.cu; 'x3' = split
    fp16, fp32-grade: BASELINE config 2's 1e-4); 'cuda' materialises the grouped tensor (rfd_query_and_group) and runs
    the fp32 CUDA-core layer kernel (the round-1 path, kept as the exact yardstick)."""

    def __init__(self, *, mlp: List[int], npoint: int = None, radius: float = None, nsample: int = None,
                 bn: bool = True, use_xyz: bool = True, pooling: str = 'max', sigma: float = None,
                 normalize_xyz: bool = False, sample_uniformly: bool = False, ret_unique_cnt: bool = False,
                 precision: str = 'x3'):
        super().__init__()
        self.precision = precision
        self.npoint, self.radius, self.nsample = npoint, radius, nsample
        self.pooling = pooling
        self.use_xyz = use_xyz
        self.sigma = sigma if sigma is not None else (self.radius / 2 if self.radius is not None else None)
        self.normalize_xyz = normalize_xyz
        self.ret_unique_cnt = ret_unique_cnt
        self.sample_uniformly = sample_uniformly
        if npoint is not None:
            self.grouper = pointnet2_utils.QueryAndGroup(radius, nsample, use_xyz=use_xyz, ret_grouped_xyz=True,
                                                         normalize_xyz=normalize_xyz,
                                                         sample_uniformly=sample_uniformly,
                                                         ret_unique_cnt=ret_unique_cnt)
        else:
            self.grouper = pointnet2_utils.GroupAll(use_xyz, ret_grouped_xyz=True)
        mlp_spec = list(mlp)
        if use_xyz and len(mlp_spec) > 0:
            mlp_spec[0] += 3
        self.mlp_module = build_shared_mlp(mlp_spec, bn=bn)
        self._fold = None

    def _fast_ok(self, xyz, features):
        return (not self.training and not torch.is_grad_enabled() and self.npoint is not None
                and self.pooling == 'max' and not self.sample_uniformly and xyz.is_cuda
                and xyz.dtype == torch.float32 and (features is None or (features.dtype == torch.float32
                                                                         and features.device == xyz.device))
                and self.mlp_module[0].weight.device == xyz.device
                and self.nsample in (4, 8, 16, 32, 64, 128))

    def forward(self, xyz, features=None, inds=None, new_xyz=None):
        """`new_xyz` (B,npoint,3), optional and only honoured together with `inds` on the fused inference path: the
        sampled coordinates when the caller already ran FPS."""
        if self._fast_ok(xyz, features):
            return self._forward_fused(xyz, features, inds, new_xyz)[:3]
        return self._forward_reference(xyz, features, inds)

    def _chain(self, layers):
        """tensor-core chain of this layer's shared MLP (None: precision 'cuda' or unsupported widths)"""
        if self.precision == 'cuda' or not self.use_xyz or len(layers) > 3 or self.nsample not in (16, 32, 64, 128):
            return None
        tc = getattr(self, "_tc", None)
        if tc is None or tc[0] != self.precision:
            tc = self._tc = (self.precision, _mlp.ChainMlp(layers, xyz=3, mode=self.precision))
        return tc[1] if tc[1].ok else None

    # -- inference: sm_100a kernels end to end
    def _forward_fused(self, xyz, features, inds, new_xyz=None, features_pm=None, want_pm=False):
        """-> (new_xyz, new_features (B,C,npoint), inds, new_features point-major (B,npoint,C) | None).
        features_pm: the same features already in point-major layout (B,N,C) (the previous layer's second output)."""
        xyz = xyz.contiguous()
        B, N, _ = xyz.shape
        if inds is not None and new_xyz is not None:
            assert inds.shape[1] == self.npoint and new_xyz.shape[1] == self.npoint
        elif inds is None:
            inds, new_xyz = pointnet2_utils.fps_with_xyz(xyz, self.npoint)  # coordinates come out of the FPS kernel
        else:
            assert inds.shape[1] == self.npoint
            new_xyz = pointnet2_utils._ext.gather_points(xyz.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()
        layers = self._folded(self.mlp_module)
        chain = self._chain(layers)
        if chain is not None:
            idx = pointnet2_utils._ext.ball_query(new_xyz, xyz, self.radius, self.nsample)
            if features is None:
                fpm = None
            elif features_pm is not None:
                fpm = features_pm
            elif features.shape[1] == 1:
                fpm = features.contiguous().view(B, N, 1)  # one channel: channel-major == point-major
            else:
                fpm = _mlp.transpose_to_point_major(features.contiguous())
                if fpm.shape[2] != features.shape[1]:
                    fpm = fpm[:, :, :features.shape[1]].contiguous()
            cm, pm = chain.gather(xyz, new_xyz, fpm, idx, self.radius, self.normalize_xyz, True, want_pm)
            return new_xyz, cm, inds, pm
        grouped, _, _ = pointnet2_utils.fused_query_and_group(
            xyz, new_xyz, None if features is None else features.contiguous(), self.radius, self.nsample,
            self.use_xyz, self.normalize_xyz)
        Ct = grouped.shape[1]
        x = grouped.view(B, Ct, self.npoint * self.nsample)
        new_features = _mlp.run_mlp(x, layers, pool_last=self.nsample)
        return new_xyz, new_features, inds, None

    # -- training / generic: the reference's sequence (:219-260) on the drop-in ops
    def _forward_reference(self, xyz, features, inds):
        xyz_flipped = xyz.transpose(1, 2).contiguous()
        if inds is None:
            inds = pointnet2_utils.furthest_point_sample(xyz, self.npoint) if self.npoint is not None else None
        else:
            assert inds.shape[1] == self.npoint
        new_xyz = pointnet2_utils.gather_operation(xyz_flipped, inds).transpose(1, 2).contiguous() \
            if self.npoint is not None else None
        if not self.ret_unique_cnt:
            grouped_features, grouped_xyz = self.grouper(xyz, new_xyz, features)
        else:
            grouped_features, grouped_xyz, unique_cnt = self.grouper(xyz, new_xyz, features)
        new_features = self.mlp_module(grouped_features)
        if self.pooling == 'max':
            new_features = F.max_pool2d(new_features, kernel_size=[1, new_features.size(3)])
        elif self.pooling == 'avg':
            new_features = F.avg_pool2d(new_features, kernel_size=[1, new_features.size(3)])
        elif self.pooling == 'rbf':
            rbf = torch.exp(-1 * grouped_xyz.pow(2).sum(1, keepdim=False) / (self.sigma ** 2) / 2)
            new_features = torch.sum(new_features * rbf.unsqueeze(1), -1, keepdim=True) / float(self.nsample)
        new_features = new_features.squeeze(-1)
        if not self.ret_unique_cnt:
            return new_xyz, new_features, inds
        return new_xyz, new_features, inds, unique_cnt


class PointnetFPModule(_FoldCache, nn.Module):
    """pointnet2_modules.py:345-405."""

    def __init__(self, mlp, bn=True, precision='x3'):
        super().__init__()
        self.mlp = build_shared_mlp(list(mlp), bn=bn)
        self.precision = precision  # 'x3' | 'fp16' | 'bf16' tcgen05 chain, 'cuda' fp32 CUDA-core layers
        self._fold = None

    def forward(self, unknown, known, unknow_feats, known_feats):
        fast = (not self.training and not torch.is_grad_enabled() and known is not None and unknown.is_cuda
                and unknown.dtype == known.dtype == known_feats.dtype == torch.float32
                and (unknow_feats is None or unknow_feats.dtype == torch.float32)
                and self.mlp[0].weight.device == unknown.device)
        if fast:
            unknown, known = unknown.contiguous(), known.contiguous()
            known_feats = known_feats.contiguous()
            B, n, _ = unknown.shape
            m = known.shape[1]
            C2 = known_feats.shape[1]
            C1 = 0 if unknow_feats is None else unknow_feats.shape[1]
            x = torch.empty((B, C2 + C1, n), dtype=torch.float32, device=unknown.device)
            with torch.cuda.device(unknown.device):
                _lib.check(_lib.load().rfd_three_nn_interpolate(
                    unknown.data_ptr(), known.data_ptr(), known_feats.data_ptr(), B, n, m, C2, C2 + C1,
                    x.data_ptr(), torch.cuda.current_stream().cuda_stream), "three_nn_interpolate")
            if C1:
                x[:, C2:, :] = unknow_feats
            layers = self._folded(self.mlp)
            if self.precision != 'cuda' and len(layers) <= 3:
                tc = getattr(self, "_tc", None)
                if tc is None or tc[0] != self.precision:
                    tc = self._tc = (self.precision, _mlp.ChainMlp(layers, xyz=0, mode=self.precision))
                if tc[1].ok:
                    return tc[1].dense(x)[0]
            return _mlp.run_mlp(x, layers)
        if known is not None:
            dist, idx = pointnet2_utils.three_nn(unknown, known)
            dist_recip = 1.0 / (dist + 1e-8)
            norm = torch.sum(dist_recip, dim=2, keepdim=True)
            weight = dist_recip / norm
            interpolated_feats = pointnet2_utils.three_interpolate(known_feats, idx, weight)
        else:
            interpolated_feats = known_feats.expand(*(list(known_feats.size()[0:2]) + [unknown.size(1)]))
        if unknow_feats is not None:
            new_features = torch.cat([interpolated_feats, unknow_feats], dim=1)
        else:
            new_features = interpolated_feats
        new_features = new_features.unsqueeze(-1)
        new_features = self.mlp(new_features)
        return new_features.squeeze(-1)
