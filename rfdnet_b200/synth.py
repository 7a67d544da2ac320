"""Seeded synthetic inputs shaped like the reference's data: ScanNet-like rooms
(points on the floor, four walls and random boxes of a 7 x 7 x 2.5 m room, 5 mm jitter, height feature
= z - percentile_0.99(z) as in models/iscnet/dataloader.py:78-81), uniform clouds, and seeded weights."""
import numpy as np
import torch


def uniform_cloud(B, N, seed=0, lo=-1.0, hi=1.0):
    rng = np.random.default_rng(seed)
    return rng.uniform(lo, hi, (B, N, 3)).astype(np.float32)


def tricky_cloud(N=4096, seed=0, n_dup=64, n_origin=4):
    """Uniform cloud with exact duplicate points (FPS / 3-NN tie-breaks) and points inside |p|^2 <= 1e-3
    (the FPS skip rule, sampling_gpu.cu:100-101)."""
    rng = np.random.default_rng(seed)
    p = rng.uniform(-1, 1, (N, 3)).astype(np.float32)
    src = rng.choice(N, n_dup, replace=False)
    dst = rng.choice(N, n_dup, replace=False)
    p[dst] = p[src]
    org = rng.choice(N, n_origin, replace=False)
    p[org] = rng.uniform(-0.015, 0.015, (n_origin, 3)).astype(np.float32)
    return p[None]


def scannet_like_scene(N=80000, seed=0):
    """(N,4) float32: xyz + height."""
    rng = np.random.default_rng(seed)
    W, D, H = 7.0, 7.0, 2.5
    surfaces = [("floor", W * D)] + [("wall%d" % i, (W if i < 2 else D) * H) for i in range(4)]
    boxes = []
    for _ in range(20):
        s = rng.uniform(0.3, 2.0, 3) * np.array([1, 1, 0.6])
        c = np.array([rng.uniform(-W / 2 + 1, W / 2 - 1), rng.uniform(-D / 2 + 1, D / 2 - 1), s[2] / 2])
        boxes.append((c, s))
        surfaces.append(("box", 2 * (s[0] * s[1] + s[0] * s[2] + s[1] * s[2])))
    areas = np.array([a for _, a in surfaces])
    counts = rng.multinomial(N, areas / areas.sum())
    pts = []
    for (name, _), n in zip(surfaces[:5], counts[:5]):
        u, v = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
        if name == "floor":
            q = np.stack([(u - .5) * W, (v - .5) * D, np.zeros(n)], 1)
        else:
            i = int(name[-1])
            if i < 2:
                q = np.stack([(u - .5) * W, np.full(n, (i - .5) * D), v * H], 1)
            else:
                q = np.stack([np.full(n, (i - 2.5) * W), (u - .5) * D, v * H], 1)
        pts.append(q)
    for (c, s), n in zip(boxes, counts[5:]):
        face = rng.integers(0, 6, n)
        q = rng.uniform(-.5, .5, (n, 3))
        ax = face % 3
        q[np.arange(n), ax] = np.where(face < 3, -.5, .5)
        pts.append(c + q * s)
    p = np.concatenate(pts, 0)
    p = p + rng.normal(0, 0.005, p.shape)
    p = p[rng.permutation(len(p))][:N].astype(np.float32)
    height = p[:, 2] - np.percentile(p[:, 2], 0.99)
    return np.concatenate([p, height[:, None].astype(np.float32)], 1).astype(np.float32)


def scannet_like_batch(B, N=80000, seed0=0):
    return np.stack([scannet_like_scene(N, seed0 + i) for i in range(B)], 0)


def seeded_fill(module_or_sd, seed=0, scale=None):
    """Deterministically (re)initialise EVERY tensor of a state_dict in key order -- including the tensors the
    reference zero-initialises (fc_1.weight, CBN conv weights: layers.py:96,220-224) and BN running statistics --
    so that parity tests cannot pass by accident.  Works on any nn.Module or plain dict with identical keys."""
    sd = module_or_sd if isinstance(module_or_sd, dict) else module_or_sd.state_dict()
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for key in sorted(sd.keys()):
            t = sd[key]
            k = "." + key
            if k.endswith("num_batches_tracked"):
                continue
            if k.endswith("running_var"):
                v = torch.rand(t.shape, generator=g) + 0.5
            elif k.endswith("running_mean"):
                v = torch.randn(t.shape, generator=g) * 0.1
            elif k.endswith(".bias"):
                v = torch.randn(t.shape, generator=g) * 0.1
            elif t.dim() == 1 and k.endswith(".weight"):
                v = torch.rand(t.shape, generator=g) + 0.5          # BN affine weight
            elif "conv_gamma.weight" in k or "conv_beta.weight" in k:
                fan_in = t[0].numel()
                v = torch.randn(t.shape, generator=g) * (0.5 / fan_in ** 0.5)
            else:
                fan_in = max(1, t[0].numel()) if t.dim() > 1 else 1
                v = torch.randn(t.shape, generator=g) * ((scale or 1.0) / fan_in ** 0.5)
            t.copy_(v.to(t.dtype))
    if not isinstance(module_or_sd, dict):
        module_or_sd.load_state_dict(sd)
    return sd
