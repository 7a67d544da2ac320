"""CPU tests of the oracle: committed golden vectors + hand-made known answers for the reference semantics
(tie-break, skip rule, empty ball, padding, m > n).  Citations: _ext-src/src/*.cu in /root/reference."""
import numpy as np
import pytest

import oracle
from rfdnet_b200.synth import tricky_cloud, uniform_cloud


def _gather(cloud, idx):
    return np.take_along_axis(cloud, idx[..., None].astype(np.int64).repeat(3, -1), 1)


@pytest.mark.parametrize("tag,cloud", [("uniform", uniform_cloud(2, 4096, seed=0)), ("tricky", tricky_cloud(4096, seed=1))])
def test_golden_config1(golden, tag, cloud):
    fps = oracle.furthest_point_sampling(cloud, 512)
    assert np.array_equal(fps, golden[f"c1_{tag}_fps"])
    q = _gather(cloud, fps)
    assert np.array_equal(oracle.ball_query(q, cloud, 0.2, 32), golden[f"c1_{tag}_bq02"])
    assert np.array_equal(oracle.ball_query(q, cloud, 0.4, 32), golden[f"c1_{tag}_bq04"])
    d2, nn = oracle.three_nn(cloud[:, :700], q)
    assert np.array_equal(nn, golden[f"c1_{tag}_nn_idx"])
    assert np.array_equal(d2, golden[f"c1_{tag}_nn_d2"])


def test_opt_n_threads():
    # cuda_utils.h:15-19
    expect = {1: 1, 2: 2, 3: 2, 4: 4, 8: 8, 10: 8, 16: 16, 100: 64, 256: 256, 511: 256, 512: 512, 1000: 512, 80000: 512}
    for k, v in expect.items():
        assert oracle.opt_n_threads(k) == v


def test_fps_first_is_zero_and_skip_rule():
    # sampling_gpu.cu:85-86 idx[0] = 0 ; :100-101 points with |p|^2 <= 1e-3 are never selected (except as index 0)
    p = np.array([[[0, 0, 0], [0.01, 0.01, 0.01], [1, 0, 0], [0, 2, 0], [0.02, 0, 0]]], np.float32)
    idx = oracle.furthest_point_sampling(p, 4)
    assert idx[0, 0] == 0
    assert 1 not in idx[0, 1:] and 4 not in idx[0, 1:]
    assert list(idx[0, :3]) == [0, 3, 2]
    # once every candidate has distance 0 the tie-break decides: smallest bit-reversed slot wins (slots 2 vs 3,
    # bs = 4: bitrev2(2) = 1 < bitrev2(3) = 3) -> index 2
    assert idx[0, 3] == 2


def test_fps_all_skipped_returns_zero():
    p = np.zeros((1, 16, 3), np.float32)
    assert np.array_equal(oracle.furthest_point_sampling(p, 5), np.zeros((1, 5), np.int32))


def test_fps_tie_break_bitreversal():
    # N = 1024 -> block 512.  Points 1 and 256 are equidistant maxima: the tree (sampling_gpu.cu:115-168, __update
    # :59-65 keeps the lower slot on ties) lets slot 256 (bitrev9 = 1) beat slot 1 (bitrev9 = 256).  SURVEY.md A4.
    p = np.full((1, 1024, 3), 0.5, np.float32)
    p[0, 0] = (0.5, 0.5, 0.5)
    p[0, 1] = (3.5, 0.5, 0.5)
    p[0, 256] = (-2.5, 0.5, 0.5)
    idx = oracle.furthest_point_sampling(p, 2)
    assert idx[0, 1] == 256
    # same slot (k mod 512 equal): the smaller k wins (strict '>' at :108-109)
    p[0, 256] = (0.5, 0.5, 0.5)
    p[0, 513] = (-2.5, 0.5, 0.5)
    assert oracle.furthest_point_sampling(p, 2)[0, 1] == 1


def test_fps_m_greater_than_n():
    p = uniform_cloud(1, 8, seed=3)
    idx = oracle.furthest_point_sampling(p, 12)
    assert idx.shape == (1, 12) and idx.min() >= 0 and idx.max() < 8
    assert len(set(idx[0, :8].tolist())) == 8  # the first n picks are all distinct


def test_ball_query_semantics():
    xyz = np.array([[[0, 0, 0], [0.05, 0, 0], [5, 5, 5], [0.0, 0.08, 0], [0.1, 0, 0]]], np.float32)
    q = np.array([[[0, 0, 0], [9, 9, 9], [5, 5, 5]]], np.float32)
    idx = oracle.ball_query(q, xyz, 0.1, 4)
    # ascending index order, strict d2 < r2 (point 4 at distance exactly r is excluded), pad with the first hit
    assert idx[0, 0].tolist() == [0, 1, 3, 0]
    # no neighbour at all: the row keeps torch::zeros (ball_query.cpp:19-21)
    assert idx[0, 1].tolist() == [0, 0, 0, 0]
    # single hit fills every slot (ball_query_gpu.cu:35-39)
    assert idx[0, 2].tolist() == [2, 2, 2, 2]
    # more hits than nsample: only the first nsample by index
    assert oracle.ball_query(q, xyz, 0.1, 2)[0, 0].tolist() == [0, 1]


def test_three_nn_ties_and_few_known():
    known = np.array([[[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0]]], np.float32)
    unk = np.zeros((1, 1, 3), np.float32)
    d2, idx = oracle.three_nn(unk, known)
    assert idx[0, 0].tolist() == [0, 1, 2]  # all equidistant: earliest indices (strict '<')
    assert np.allclose(d2, 1.0)
    d2, idx = oracle.three_nn(unk, known[:, :2])  # fewer than 3 known: 1e40 -> +inf, index 0
    assert idx[0, 0].tolist() == [0, 1, 0] and np.isinf(d2[0, 0, 2])


def test_group_gather_interp_and_grads():
    rng = np.random.default_rng(0)
    pts = rng.normal(size=(2, 5, 50)).astype(np.float32)
    idx = rng.integers(0, 50, (2, 7, 4)).astype(np.int32)
    g = oracle.group_points(pts, idx)
    assert g.shape == (2, 5, 7, 4)
    assert np.array_equal(g[1, 3, 2, 1], pts[1, 3, idx[1, 2, 1]])
    go = rng.normal(size=g.shape).astype(np.float32)
    gg = oracle.group_points_grad(go, idx, 50)
    ref = np.zeros((2, 5, 50), np.float64)
    for b in range(2):
        for j in range(7):
            for k in range(4):
                ref[b, :, idx[b, j, k]] += go[b, :, j, k]
    assert np.allclose(gg, ref, atol=1e-5)
    i1 = rng.integers(0, 50, (2, 9)).astype(np.int32)
    ga = oracle.gather_points(pts, i1)
    assert np.array_equal(ga[0, :, 4], pts[0, :, i1[0, 4]])
    w = rng.uniform(size=(2, 9, 3)).astype(np.float32)
    i3 = rng.integers(0, 50, (2, 9, 3)).astype(np.int32)
    out = oracle.three_interpolate(pts, i3, w)
    ref = sum(np.take_along_axis(pts, i3[:, None, :, t].repeat(5, 1), 2) * w[:, None, :, t] for t in range(3))
    assert np.allclose(out, ref, atol=1e-6)
    gi = oracle.three_interpolate_grad(out, i3, w, 50)
    assert gi.shape == (2, 5, 50)
    assert np.isclose(gi.sum(), (out[:, :, :, None] * w[:, None]).sum(), rtol=1e-4)


def test_fps_rank_formula_matches_tree():
    """The product kernel replaces the shared-memory tree by a max over the key (value, -rank) with
    rank(k) = bitrev(k mod bs) * ceil(N/bs) + k div bs.  Check the closed form against the literal tree of the
    oracle on clouds FULL of exact ties (lattice points)."""
    rng = np.random.default_rng(5)
    for N in (37, 600, 1500, 4096):
        p = rng.integers(-3, 4, (1, N, 3)).astype(np.float32)  # heavy duplication => many ties
        m = min(N, 40)
        got = oracle.furthest_point_sampling(p, m)[0]
        bs = oracle.opt_n_threads(N)
        lg = bs.bit_length() - 1
        Q = -(-N // bs)
        x = p[0]
        mag = (x[:, 1] * x[:, 1]).astype(np.float32)
        mag = (x[:, 0].astype(np.float64) * x[:, 0] + mag).astype(np.float32)  # exact for small integers
        mag = (x[:, 2].astype(np.float64) * x[:, 2] + mag).astype(np.float32)
        ok = mag.astype(np.float64) > 1e-3
        k = np.arange(N)
        slot = k % bs
        rev = np.array([int(format(s, f"0{lg}b")[::-1], 2) if lg else 0 for s in slot])
        rank = rev * Q + k // bs
        temp = np.full(N, 1e10, np.float32)
        sel = [0]
        old = 0
        for _ in range(1, m):
            d = ((x - x[old]) ** 2).sum(1).astype(np.float32)
            temp = np.where(ok, np.minimum(d, temp), temp)
            cand = np.where(ok, temp, -1.0)
            best = cand.max()
            if best < 0:
                old = 0
            else:
                tied = np.where(cand == best)[0]
                old = int(tied[np.argmin(rank[tied])])
            sel.append(old)
        assert got.tolist() == sel, N
