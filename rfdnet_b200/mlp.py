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
    check_f32(x, "x")
    B, Cin, L = x.shape
    Cout = W.shape[0]
    assert W.shape[1] == Cin and W.device == x.device, (W.shape, Cin)
    if pool > 64 and pool % 64 == 0:
        # the kernel pools inside its 64-position tile; wider groups (STN_Group-sized nsample) finish with one reduction
        # over the per-tile maxima
        y = pointwise_layer(x, W, scale, shift, relu, pool=64, residual=residual)
        return y.view(B, Cout, L // pool, pool // 64).amax(dim=-1)
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


MODES = {'bf16': 1, 'fp16': 2, 'x3': 3}


def _ptr(t):
    return 0 if t is None else t.data_ptr()


class ChainMlp:
    """1-3 folded pointwise layers [(W, scale, shift, relu), ...] packed for the tcgen05 chain kernel
    (csrc/mlp_chain_tc.cu).  mode: 'x3' (split-fp16, fp32-grade: BASELINE config 2's 1e-4), 'fp16', 'bf16'.
    xyz = 3: the first three input columns of layer 0 are the relative-xyz channels of an SA layer, applied in fp32
    by the epilogue (gather mode).  `ok` is False when the widths are outside what the kernel supports."""

    def __init__(self, layers, xyz=0, mode='x3'):
        n = len(layers)
        assert 1 <= n <= 3 and all(l[3] for l in layers[:-1]), "intermediate layers must have a ReLU"
        self.mode, self.xyz = MODES[mode], int(xyz)
        self.relu_last = int(bool(layers[-1][3]))
        self.K0 = layers[0][0].shape[1] - self.xyz
        self.C = [l[0].shape[0] for l in layers] + [0] * (3 - n)
        self.out_C = layers[-1][0].shape[0]
        # padded width of layer 0's accumulator (the row length of a gbias table)
        self.n0 = (self.C[0] + 63) // 64 * 64 if n > 1 else (min(self.C[0], 256) + 15) // 16 * 16
        self.flop_per_row = 2.0 * sum(l[0].shape[0] * l[0].shape[1] for l in layers)
        lib = _lib.load()
        nbytes = lib.rfd_mlp_chain_packed_bytes(self.mode, self.K0, self.xyz, *self.C)
        self.ok = nbytes > 0
        if not self.ok:
            return
        dev = layers[0][0].device
        self.packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        args = []
        for i in range(3):
            if i < n:
                W, s, t, _ = layers[i]
                self._keep = getattr(self, "_keep", []) + [W.contiguous(), s.contiguous(), t.contiguous()]
                args += [self._keep[-3].data_ptr(), self._keep[-2].data_ptr(), self._keep[-1].data_ptr(), self.C[i]]
            else:
                args += [0, 0, 0, 0]
        with torch.cuda.device(dev):
            _lib.check(lib.rfd_mlp_chain_pack(self.mode, self.K0, self.xyz, *args, self.relu_last,
                                              self.packed.data_ptr(), torch.cuda.current_stream().cuda_stream),
                       "mlp_chain_pack")
        self._keep = None  # the packed image owns copies of everything

    def _outs(self, B, rows, dev, want_cm, want_pm):
        cm = torch.empty((B, self.out_C, rows), dtype=torch.float32, device=dev) if want_cm else None
        pm = torch.empty((B, rows, self.out_C), dtype=torch.float32, device=dev) if want_pm else None
        return cm, pm

    def dense(self, x, pool=1, want_cm=True, want_pm=False, relu_in=False, gbias=None, gbias_rows=0, out_cm=None,
              out_pool=None, pool_rows=0):
        """x (B, K0, L) f32 channel-major -> (out_cm (B, C, L/pool) | None, out_pm (B, L/pool, C) | None).
        relu_in / gbias (B, L/gbias_rows, n0) / out_pool (B, C, L/pool_rows, caller-initialised to -inf): see
        rfd_mlp_chain_ex.  out_cm: write into this preallocated (B, C, L/pool) contiguous view instead of a new tensor."""
        check_f32(x, "x")
        B, K0, L = x.shape
        assert K0 == self.K0 and self.xyz == 0, (K0, self.K0)
        cm, pm = self._outs(B, L // pool, x.device, want_cm and out_cm is None, want_pm)
        if out_cm is not None:
            assert out_cm.shape == (B, self.out_C, L // pool) and out_cm.is_contiguous() and out_cm.dtype == torch.float32
            cm = out_cm
        if gbias is not None:
            check_f32(gbias, "gbias")
            assert gbias.shape[0] == B and gbias.shape[1] == L // gbias_rows and gbias.shape[2] == self.n0, (gbias.shape, self.n0)
        if out_pool is not None:
            assert out_pool.shape == (B, self.out_C, L // pool_rows) and out_pool.is_contiguous()
        with torch.cuda.device(x.device), _lib.timed("mlp_chain_tc", B * L * self.flop_per_row):
            _lib.check(_lib.load().rfd_mlp_chain_ex(self.mode, x.data_ptr(), B, K0, L, self.packed.data_ptr(), *self.C,
                                                    self.relu_last, int(pool), _ptr(cm), _ptr(pm), int(bool(relu_in)),
                                                    _ptr(gbias), int(gbias_rows), _ptr(out_pool), int(pool_rows),
                                                    torch.cuda.current_stream().cuda_stream), "mlp_chain")
        return cm, pm

    def rows(self, x, out=None, out_col0=0, want_out=True, relu_in=False, gbias=None, gbias_rows=0, out_pool=None, pool_rows=0):
        """Row-major operands (rfd_mlp_chain_rows): x (B, L, ldi) f32, the K0 operand channels in columns [0, K0) of every
        row -> out (B, L, ldo) columns [out_col0, out_col0 + C) (given, or a new (B, L, C) tensor; want_out=False: none);
        gbias (B, L/gbias_rows, n0); out_pool (B, L/pool_rows, C), caller-initialised to -inf."""
        check_f32(x, "x")
        B, L, ldi = x.shape
        assert ldi >= self.K0 and self.xyz == 0 and x.stride(1) == ldi, (x.shape, x.stride(), self.K0)
        if out is None and want_out:
            out = torch.empty((B, L, self.out_C), dtype=torch.float32, device=x.device)
            out_col0 = 0
        ldo = 0
        if out is not None:
            assert out.shape[:2] == (B, L) and out.stride(1) == out.shape[2] and out.dtype == torch.float32
            ldo = out.shape[2]
        if gbias is not None:
            check_f32(gbias, "gbias")
            assert gbias.shape == (B, L // gbias_rows, self.n0), (gbias.shape, self.n0)
        if out_pool is not None:
            assert out_pool.shape == (B, L // pool_rows, self.out_C) and out_pool.is_contiguous()
        with torch.cuda.device(x.device), _lib.timed("mlp_chain_tc", B * L * self.flop_per_row):
            _lib.check(_lib.load().rfd_mlp_chain_rows(self.mode, x.data_ptr(), ldi, B, self.K0, L, self.packed.data_ptr(),
                                                      *self.C, self.relu_last, _ptr(out), ldo, int(out_col0),
                                                      int(bool(relu_in)), _ptr(gbias), int(gbias_rows), _ptr(out_pool),
                                                      int(pool_rows), torch.cuda.current_stream().cuda_stream),
                       "mlp_chain_rows")
        return out

    def gather(self, xyz, new_xyz, feat_pm, idx, radius, normalize_xyz, want_cm=True, want_pm=False):
        """Full SA fusion: rows gathered through idx (B,M,S) from xyz (B,N,3) / point-major features (B,N,C),
        centred on new_xyz, 3-layer MLP, max over S -> (out_cm (B,C3,M) | None, out_pm (B,M,C3) | None)."""
        check_f32(xyz, "xyz"); check_f32(new_xyz, "new_xyz")
        B, N, _ = xyz.shape
        _, M, S = idx.shape
        C = 0 if feat_pm is None else feat_pm.shape[2]
        if feat_pm is not None:
            check_f32(feat_pm, "features")
            assert feat_pm.shape[:2] == (B, N)
        assert C == self.K0 and self.xyz == 3 and idx.dtype == torch.int32 and idx.is_contiguous()
        cm, pm = self._outs(B, M, xyz.device, want_cm, want_pm)
        with torch.cuda.device(xyz.device), _lib.timed("mlp_chain_tc", B * M * S * self.flop_per_row):
            _lib.check(_lib.load().rfd_sa_mlp_chain(
                self.mode, xyz.data_ptr(), new_xyz.data_ptr(), _ptr(feat_pm), idx.data_ptr(), B, N, M, S, C,
                float(radius), int(bool(normalize_xyz)), self.packed.data_ptr(), *self.C, _ptr(cm), _ptr(pm),
                torch.cuda.current_stream().cuda_stream), "sa_mlp_chain")
        return cm, pm


class WideLayer:
    """ONE pointwise layer of any input / output width on the tcgen05 chain kernel, for the PointNet-style encoders of
    SkipPropagation (widths up to 1536 -> 1024): the output channels are split into blocks of <= 256, one single-layer
    chain launch per block, every launch streaming the whole K through the resident A panels.  Operands are ROW-major --
    x (1, R, ld) with the K operand channels in the first K columns of every row -- so a tile reads 128 contiguous rows
    and channel concatenation is a column offset into a wider row."""

    def __init__(self, W, scale, shift, relu, mode='x3', block=256):
        self.Cout, self.K = W.shape
        self.blocks = []
        for c0 in range(0, self.Cout, block):
            c1 = min(self.Cout, c0 + block)
            ch = ChainMlp([(W[c0:c1].contiguous(), scale[c0:c1].contiguous(), shift[c0:c1].contiguous(), bool(relu))],
                          xyz=0, mode=mode)
            assert ch.ok, (W.shape, c0, c1)
            self.blocks.append((c0, c1, ch))

    def __call__(self, x, out=None, out_col0=0, relu_in=False, gbias=None, gbias_rows=0, pool_rows=0):
        """x (1,R,ld) -> out (1,R,ldo) columns [out_col0, out_col0 + Cout) (None = the rows are not written);
        gbias (1,G,Cout) per-group bias; pool_rows > 0: also returns the max over groups of pool_rows rows, (1,G',Cout)."""
        assert x.shape[0] == 1 and x.shape[2] >= self.K, (x.shape, self.K)
        pooled = []
        for c0, c1, ch in self.blocks:
            gb = None
            if gbias is not None:
                gb = torch.zeros((1, gbias.shape[1], ch.n0), dtype=torch.float32, device=x.device)
                gb[:, :, :c1 - c0] = gbias[:, :, c0:c1]
            pl = None
            if pool_rows:
                pl = torch.full((1, x.shape[1] // pool_rows, c1 - c0), float("-inf"), dtype=torch.float32, device=x.device)
                pooled.append(pl)
            ch.rows(x, out=out, out_col0=out_col0 + c0, want_out=False, relu_in=relu_in, gbias=gb, gbias_rows=gbias_rows,
                    out_pool=pl, pool_rows=pool_rows)
        return torch.cat(pooled, dim=2) if pool_rows else out


def state_version(module):
    """cheap fingerprint of a module's parameters and buffers: (data_ptr, in-place version counter) of each.  The folded /
    packed weight caches of the inference paths are keyed on it, so an in-place update (optimizer step, manual edit) while
    the module stays in eval mode is noticed on the next forward."""
    return tuple((t.data_ptr(), t._version) for t in list(module.parameters()) + list(module.buffers()))


def check_f32(t, name):
    """Same precondition errors as the reference's CHECK_CUDA / CHECK_CONTIGUOUS / CHECK_IS_FLOAT (utils.h:5-25): the
    raw-pointer entry points must never reinterpret a half / double / CPU / strided tensor as dense fp32."""
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be a float tensor")


def transpose_to_point_major(features):
    """(B, C, N) channel-major -> (B, N, Cp) point-major, Cp = C rounded up to 4 (rfd_transpose_features)."""
    check_f32(features, "features")
    B, C, N = features.shape
    Cp = (C + 3) & ~3
    out = torch.empty((B, N, Cp), dtype=torch.float32, device=features.device)
    with torch.cuda.device(features.device):
        _lib.check(_lib.load().rfd_transpose_features(features.data_ptr(), B, C, N, Cp, out.data_ptr(),
                                                      torch.cuda.current_stream().cuda_stream), "transpose_features")
    return out
