"""STN_Group -- the per-proposal grouping of SkipPropagation on the sm_100a kernels.

What the reference computes (external/pointnet2_ops_lib/pointnet2_ops/pointnet2_modules.py:420-537, called from
models/iscnet/modules/skip_propagation.py:84-129 with radius 1.0 / nsample 1024 over the whole 80k-point cloud):
group the scene points around every kept box centre, rotate the relative coordinates into the box frame, regress a
3x4 alignment with a small PointNet (STN3d: three 1x1 convolutions, max over the 1024 points, three linear layers) and
apply it.  The reference does this with ball_query + 2 group_points + sub + cat, a zero-initialised rotation tensor
filled by five indexed writes, two bmm, two transposes with copies, six cuDNN / cuBLAS launches and their BatchNorms.

Inference here is four launches of librfdnet_b200, nothing else:
  1. rfd_query_and_group_rotated   ball query (index-ordered scan, stops at 1024 hits), feature gather, centring and
                                   the heading rotation in ONE kernel                                (csrc/ballquery_group.cu)
  2. rfd_mlp_chain (K0 = 3 -> 64 -> 128 -> 256, max over the 1024 samples)   tcgen05 chain, folded BN  (csrc/mlp_chain_tc.cu)
  3. rfd_mlp_chain (256 -> 128 -> 64 -> 12 on the pooled rows)               tcgen05 chain, folded BN
  4. rfd_stn_apply                 (theta + [I|0]) applied to the rotated coordinates
Under autograd (training) the op sequence is spelled out on the drop-in operators so gradients flow as in the reference.
Parameter names and shapes equal the reference's, so its checkpoints load.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, mlp as _mlp, pointnet2_utils


def _linear_bn_as_layer(fc, bn):
    """nn.Linear (+ eval BatchNorm1d) as a folded pointwise layer (W, scale, shift) acting on channel-major rows"""
    W = fc.weight.detach().float().contiguous()
    if bn is None:
        return W, torch.ones_like(fc.bias.detach().float()), fc.bias.detach().float().contiguous()
    s = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    t = (fc.bias.detach().float() - bn.running_mean.detach().float()) * s + bn.bias.detach().float()
    return W, s.contiguous(), t.contiguous()


class STN3d(nn.Module):
    """Alignment regressor; `forward(g)` takes and returns (B, 3, K, S) like the reference's module."""

    def __init__(self, num_points=2500):
        super().__init__()
        self.num_points = num_points
        self.conv1, self.conv2, self.conv3 = nn.Conv1d(3, 64, 1), nn.Conv1d(64, 128, 1), nn.Conv1d(128, 256, 1)
        self.mp1 = nn.MaxPool1d(num_points)
        self.fc1, self.fc2, self.fc3 = nn.Linear(256, 128), nn.Linear(128, 64), nn.Linear(64, 12)
        self.relu = nn.ReLU(inplace=True)
        self.bn1, self.bn2, self.bn3 = nn.BatchNorm1d(64), nn.BatchNorm1d(128), nn.BatchNorm1d(256)
        self.bn4, self.bn5 = nn.BatchNorm1d(128), nn.BatchNorm1d(64)
        # the reference's weights_init (pointnet2_modules.py:407-418) matches class names containing 'Conv2d' or
        # 'Linear' only: the three Linear layers start at zero (identity alignment), the Conv1d layers keep PyTorch's init
        for m in (self.fc1, self.fc2, self.fc3):
            nn.init.zeros_(m.weight)
            nn.init.zeros_(m.bias)
        self.precision = 'x3'
        self._packed = None

    # -- folded / packed weights of the two tensor-core chains, rebuilt when a parameter or buffer changes
    def _version(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    def _chains(self):
        ver = (self._version(), self.precision)
        if self._packed is None or self._packed[0] != ver:
            convs = [(*_mlp.fold_conv_bn(c, b), True) for c, b in ((self.conv1, self.bn1), (self.conv2, self.bn2),
                                                                     (self.conv3, self.bn3))]
            fcs = [(*_linear_bn_as_layer(self.fc1, self.bn4), True), (*_linear_bn_as_layer(self.fc2, self.bn5), True),
                   (*_linear_bn_as_layer(self.fc3, None), False)]
            self._packed = (ver, _mlp.ChainMlp(convs, xyz=0, mode=self.precision), _mlp.ChainMlp(fcs, xyz=0, mode=self.precision))
        return self._packed[1], self._packed[2]

    def theta(self, g):
        """(B,3,K,S) rotated coordinates -> regressed 3x4 entries WITHOUT the identity, (B,12,K)."""
        B, _, K, S = g.shape
        conv_chain, fc_chain = self._chains()
        pooled, _ = conv_chain.dense(g.view(B, 3, K * S), pool=S)           # (B,256,K): conv x3 + max over the samples
        theta, _ = fc_chain.dense(pooled)                                   # (B,12,K)
        return theta

    def _fast_ok(self, g):
        S = g.shape[-1]
        return (not self.training and not torch.is_grad_enabled() and g.is_cuda and g.dtype == torch.float32
                and self.conv1.weight.device == g.device and S == self.num_points and S >= 64 and (S & (S - 1)) == 0)

    def forward(self, grouped_xyz):
        if self._fast_ok(grouped_xyz):
            g = grouped_xyz.contiguous()
            B, _, K, S = g.shape
            out = torch.empty_like(g)
            with torch.cuda.device(g.device):
                _lib.check(_lib.load().rfd_stn_apply(g.data_ptr(), self.theta(g).data_ptr(), B, K, S, out.data_ptr(),
                                                     torch.cuda.current_stream().cuda_stream), "stn_apply")
            return out
        # autograd path: the same function, one torch op per step
        B, _, K, S = grouped_xyz.shape
        g = grouped_xyz.permute(0, 2, 1, 3).reshape(B * K, 3, S)
        x = F.relu(self.bn1(self.conv1(g)))
        x = F.relu(self.bn2(self.conv2(x)))
        x = F.relu(self.bn3(self.conv3(x)))
        x = x.amax(dim=2)
        x = F.relu(self.bn4(self.fc1(x)))
        x = F.relu(self.bn5(self.fc2(x)))
        eye = torch.eye(3, 4, device=g.device, dtype=g.dtype).reshape(1, 12)
        T = (self.fc3(x) + eye).view(B * K, 3, 4)
        out = torch.baddbmm(T[:, :, 3:], T[:, :, :3], g)
        return out.view(B, K, 3, S).permute(0, 2, 1, 3)


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

    def _fast_ok(self, xyz, features, new_xyz, orientations):
        ts = [t for t in (xyz, features, new_xyz, orientations) if t is not None]
        return (not self.training and not torch.is_grad_enabled() and not self.grouper.sample_uniformly
                and all(t.is_cuda and t.dtype == torch.float32 and t.device == xyz.device for t in ts)
                and (features is not None or self.use_xyz) and self.nsample <= 1024)

    def forward(self, xyz, features=None, new_xyz=None, orientations=None):
        """xyz (B,N,3), features (B,C,N), new_xyz (B,K,3) box centres, orientations (B,K) headings ->
        (aligned grouped_xyz (B,3,K,S), grouped features (B,C[+3],K,S)[, unique_cnt])."""
        if self._fast_ok(xyz, features, new_xyz, orientations):
            xyz, new_xyz, heading = xyz.contiguous(), new_xyz.contiguous(), orientations.contiguous()
            feats = None if features is None else features.contiguous()
            B, N, _ = xyz.shape
            K, S = new_xyz.shape[1], self.nsample
            C = 0 if feats is None else feats.shape[1]
            grouped_features = torch.empty((B, C + (3 if self.use_xyz else 0), K, S), dtype=torch.float32, device=xyz.device)
            grouped_xyz = torch.empty((B, 3, K, S), dtype=torch.float32, device=xyz.device)
            with torch.cuda.device(xyz.device):
                _lib.check(_lib.load().rfd_query_and_group_rotated(
                    xyz.data_ptr(), new_xyz.data_ptr(), 0 if feats is None else feats.data_ptr(), heading.data_ptr(), B, N, K,
                    C, float(self.radius), int(S), int(bool(self.use_xyz)), int(bool(self.normalize_xyz)),
                    grouped_features.data_ptr(), grouped_xyz.data_ptr(), 0, torch.cuda.current_stream().cuda_stream),
                    "query_and_group_rotated")
            if self.use_xyz:
                # the reference rotates only the returned coordinates; the xyz channels inside new_features stay unrotated
                grouped_features[:, :3] = pointnet2_utils.fused_query_and_group(
                    xyz, new_xyz, None, self.radius, S, True, self.normalize_xyz)[0]
            return self.stn3d(grouped_xyz), grouped_features
        out = self.grouper(xyz, new_xyz, features)
        grouped_features, grouped_xyz = out[0], out[1]
        c, s = torch.cos(orientations), torch.sin(orientations)
        gx, gy, gz = grouped_xyz.unbind(dim=1)                           # (B,K,S) each
        grouped_xyz = torch.stack([c.unsqueeze(-1) * gx + s.unsqueeze(-1) * gy,
                                   c.unsqueeze(-1) * gy - s.unsqueeze(-1) * gx, gz], dim=1)
        grouped_xyz = self.stn3d(grouped_xyz)
        return (grouped_xyz, grouped_features, out[2]) if self.ret_unique_cnt else (grouped_xyz, grouped_features)
