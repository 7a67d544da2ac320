"""Pointwise-MLP helpers: BatchNorm folding (eval mode) and launches of the fp32 layer kernel."""
import torch
import torch.nn as nn

from . import _lib


def fold_conv_bn(conv, bn=None):
    """(W (Cout,Cin) f32, scale (Cout), shift (Cout)) with  bn(conv(x)) == scale * (W.x) + shift  in eval mode.

    conv: nn.Conv1d / nn.Conv2d with kernel size 1 (or nn.Linear); bn: nn.BatchNorm{1,2}d or None."""
    W = conv.weight.detach().reshape(conv.weight.shape[0], -1).float().contiguous()
    cout = W.shape[0]
    bias = conv.bias.detach().float() if conv.bias is not None else torch.zeros(cout, device=W.device)
    if bn is None:
        return W, torch.ones(cout, device=W.device), bias.contiguous()
    inv = torch.rsqrt(bn.running_var.detach().float() + bn.eps)
    g = bn.weight.detach().float() if bn.affine else torch.ones_like(inv)
    b = bn.bias.detach().float() if bn.affine else torch.zeros_like(inv)
    scale = g * inv
    shift = b + (bias - bn.running_mean.detach().float()) * scale
    return W, scale.contiguous(), shift.contiguous()


def fold_sequential(seq):
    """Fold an nn.Sequential of [Conv, (BN), (ReLU)]* (build_shared_mlp, pointnet2_modules.py:9-19)
    into a list of (W, scale, shift, relu)."""
    layers, mods, i = [], list(seq), 0
    while i < len(mods):
        conv = mods[i]
        assert isinstance(conv, (nn.Conv1d, nn.Conv2d)), type(conv)
        i += 1
        bn = None
        if i < len(mods) and isinstance(mods[i], (nn.BatchNorm1d, nn.BatchNorm2d)):
            bn = mods[i]
            i += 1
        relu = False
        if i < len(mods) and isinstance(mods[i], nn.ReLU):
            relu = True
            i += 1
        W, s, t = fold_conv_bn(conv, bn)
        layers.append((W, s, t, relu))
    return layers


def pointwise_layer(x, W, scale, shift, relu, pool=1, residual=None):
    """y = act(scale * (W . x) + shift) (+ residual), optional max over runs of `pool` positions.
    x (B,Cin,L) f32 contiguous CUDA -> (B,Cout,L//pool)."""
    B, Cin, L = x.shape
    Cout = W.shape[0]
    assert W.shape[1] == Cin, (W.shape, Cin)
    y = torch.empty((B, Cout, L // pool), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().rfd_pointwise_mlp_f32(
            x.data_ptr(), W.data_ptr(), scale.data_ptr(), shift.data_ptr(),
            0 if residual is None else residual.data_ptr(), int(relu), int(pool), B, Cin, Cout, L, y.data_ptr(),
            torch.cuda.current_stream().cuda_stream), "pointwise_mlp_f32")
    return y


def run_mlp(x, layers, pool_last=1):
    """Apply folded layers to x (B,C,L); the last layer max-pools over runs of `pool_last`."""
    for li, (W, s, t, relu) in enumerate(layers):
        last = li == len(layers) - 1
        x = pointwise_layer(x, W, s, t, relu, pool=pool_last if last else 1)
    return x


class PackedMlp3:
    """bf16 tensor-core weights of a 3-layer shared MLP (rfd_sa_mlp_tc_pack); None if the widths are unsupported."""

    def __init__(self, layers):
        (W1, s1, t1, r1), (W2, s2, t2, r2), (W3, s3, t3, r3) = layers
        assert r1 and r2 and r3
        self.Ct, self.C1, self.C2, self.C3 = W1.shape[1], W1.shape[0], W2.shape[0], W3.shape[0]
        lib = _lib.load()
        nbytes = lib.rfd_sa_mlp_tc_packed_bytes(self.Ct, self.C1, self.C2, self.C3)
        self.ok = nbytes > 0
        if not self.ok:
            return
        dev = W1.device
        self.packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        self.shift = torch.cat([t1, t2, t3]).contiguous()
        with torch.cuda.device(dev):
            _lib.check(lib.rfd_sa_mlp_tc_pack(W1.data_ptr(), s1.data_ptr(), W2.data_ptr(), s2.data_ptr(), W3.data_ptr(),
                                              s3.data_ptr(), self.Ct, self.C1, self.C2, self.C3, self.packed.data_ptr(),
                                              torch.cuda.current_stream().cuda_stream), "sa_mlp_tc_pack")

    def __call__(self, grouped):
        """grouped (B, Ct, M, S) f32 -> (B, C3, M) f32."""
        B, Ct, M, S = grouped.shape
        assert Ct == self.Ct
        out = torch.empty((B, self.C3, M), dtype=torch.float32, device=grouped.device)
        flop = 2.0 * B * M * S * (self.Ct * self.C1 + self.C1 * self.C2 + self.C2 * self.C3)
        with torch.cuda.device(grouped.device), _lib.timed("sa_mlp_tc", flop):
            _lib.check(_lib.load().rfd_sa_mlp_tc(grouped.data_ptr(), B, Ct, M, S, self.packed.data_ptr(),
                                                 self.shift.data_ptr(), self.C1, self.C2, self.C3, out.data_ptr(),
                                                 torch.cuda.current_stream().cuda_stream), "sa_mlp_tc")
        return out

    def fused(self, xyz, new_xyz, features, idx, radius, normalize_xyz):
        """Full SA fusion: gather + centre/normalise + 3-layer MLP + max without materialising the grouped tensor.
        xyz (B,N,3), new_xyz (B,M,3), features (B,C,N) or None, idx (B,M,S) i32 -> (B, C3, M)."""
        B, N, _ = xyz.shape
        _, M, S = idx.shape
        C = 0 if features is None else features.shape[1]
        assert 3 + C == self.Ct
        out = torch.empty((B, self.C3, M), dtype=torch.float32, device=xyz.device)
        flop = 2.0 * B * M * S * (self.Ct * self.C1 + self.C1 * self.C2 + self.C2 * self.C3)
        with torch.cuda.device(xyz.device), _lib.timed("sa_mlp_tc", flop):
            _lib.check(_lib.load().rfd_sa_gather_mlp_tc(
                xyz.data_ptr(), new_xyz.data_ptr(), 0 if features is None else features.data_ptr(), idx.data_ptr(),
                B, N, M, S, C, float(radius), int(bool(normalize_xyz)), self.packed.data_ptr(), self.shift.data_ptr(),
                self.C1, self.C2, self.C3, out.data_ptr(), torch.cuda.current_stream().cuda_stream),
                "sa_gather_mlp_tc")
        return out
