"""Surface extraction, CPU side: the marching-cubes table and the oracle restatement of
Generator3D.extract_mesh (generator.py:145-168 over PyMCubes 0.1.2, which is absent from this image).  The table is
checked by its own invariants -- nothing here can pass by copying a wrong table twice."""
import collections
import os
import re

import numpy as np

from mesh_checks import assert_closed_oriented
from oracle import cpu_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CORNERS = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
EDGES = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]


def _parse_table(path):
    txt = open(path).read()
    body = txt[txt.index("MC_TRI_TABLE[256][16]"):]
    rows = re.findall(r"\{([^{}]*)\}", body)
    return np.array([[int(x) for x in r.split(",")] for r in rows[:256]], dtype=np.int8)


def test_product_and_oracle_tables_agree():
    a = _parse_table(os.path.join(ROOT, "rfdnet_b200", "csrc", "mc_tables.h"))
    b = _parse_table(os.path.join(ROOT, "oracle", "mc_tables_oracle.h"))
    assert a.shape == (256, 16) and np.array_equal(a, b) and np.array_equal(a, cpu_ref.mc_table())


def test_table_rows_use_exactly_the_sign_change_edges():
    T = cpu_ref.mc_table()
    for c in range(256):
        row = [int(e) for e in T[c] if e >= 0]
        assert len(row) % 3 == 0 and len(row) <= 15
        assert all(e == -1 for e in T[c][len(row):])
        crossing = {e for e, (a, b) in enumerate(EDGES) if ((c >> a) & 1) != ((c >> b) & 1)}
        assert set(row) == crossing, c
        for i in range(0, len(row), 3):
            assert len(set(row[i:i + 3])) == 3
    assert (T[0] == -1).all() and (T[255] == -1).all()


def test_random_fields_give_closed_oriented_surfaces():
    rng = np.random.default_rng(0)
    seen = set()
    for trial in range(4):
        occ = rng.normal(size=(9, 9, 9)).astype(np.float32)                # white noise: every one of the 256 cases occurs
        v, t, keys = cpu_ref.extract_mesh(occ)
        assert len(np.unique(keys)) == len(keys) == len(v)
        assert t.min() >= 0 and t.max() == len(v) - 1
        assert_closed_oriented(t)
        s = np.pad(occ, 1, constant_values=-1e6) <= 0.0
        for i in range(10):
            for j in range(10):
                for k in range(10):
                    seen.add(sum(int(s[i + a, j + b, k + c]) << m for m, (a, b, c) in enumerate(CORNERS)))
    assert len(seen) == 256


def test_known_answers():
    occ = np.full((4, 4, 4), -5.0, np.float32)
    v, t, _ = cpu_ref.extract_mesh(occ)                                    # nothing occupied: empty mesh
    assert v.shape == (0, 3) and t.shape == (0, 3)
    occ[1, 2, 1] = 3.0                                                     # one occupied sample: an octahedron
    v, t, _ = cpu_ref.extract_mesh(occ)
    assert v.shape == (6, 3) and t.shape == (8, 3)
    assert_closed_oriented(t)
    # crossing at (0 - (-5)) / (3 - (-5)) = 5/8 of the way towards the occupied sample; reference transform:
    # box * ((idx_padded - 0.5 - 1) / (n-1) - 0.5), idx_padded = idx + 1
    def box(c):
        return 1.1 * ((c + 1 - 1.5) / 3 - 0.5)
    xs = sorted(v[:, 0])
    assert np.allclose([xs[0], xs[-1]], [box(1 - 3 / 8), box(1 + 3 / 8)], atol=1e-12)
    centre = v.mean(0)
    assert np.allclose(centre, [box(1), box(2), box(1)], atol=1e-12)
    # outward orientation: signed volume positive or negative consistently; mcubes' convention gives a closed surface whose
    # signed volume has one sign for one blob -- check it is non-zero and equals the octahedron volume in magnitude
    a, b, c = v[t[:, 0]], v[t[:, 1]], v[t[:, 2]]
    vol = np.einsum("ij,ij->i", a, np.cross(b, c)).sum() / 6
    r = 1.1 * (3 / 8) / 3
    assert np.isclose(abs(vol), 4 / 3 * r ** 3, rtol=1e-9)
    occ[:] = 7.0                                                            # everything occupied: the padding closes a box
    v, t, _ = cpu_ref.extract_mesh(occ)
    assert_closed_oriented(t)
    assert len(v) == 6 * 16 and len(t) == 2 * len(v) - 4                    # genus-0 closed surface: T = 2V - 4
    assert np.isclose(v.min(), box(0 - 1 + 1e6 / (1e6 + 7)), atol=1e-9)


def test_equal_values_and_threshold_ties():
    occ = np.zeros((3, 3, 3), np.float32)                                   # value == threshold counts as "outside" (<=)
    v, t, _ = cpu_ref.extract_mesh(occ)
    assert len(v) == 0 and len(t) == 0
    occ[1, 1, 1] = np.nextafter(np.float32(0), np.float32(1))
    v, t, _ = cpu_ref.extract_mesh(occ)
    assert len(v) == 6 and len(t) == 8


def test_oracle_matches_committed_golden_meshes():
    """tests/golden/golden_mesh.npz (made by tests/golden/make_golden_mesh.py) pins the oracle restatement itself."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_mesh", os.path.join(ROOT, "tests", "golden", "make_golden_mesh.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "golden_mesh.npz"))
    for name, f in mg.fields().items():
        for thr, pad in ((0.5, 0.1), (0.3, 0.25)):
            d = mg.digest(*cpu_ref.extract_mesh(f, thr, pad))
            assert d["nv"] > 0
            for k, val in d.items():
                g = gold[f"{name}_{thr}_{pad}_{k}"]
                assert np.array_equal(np.asarray(val), g), (name, thr, pad, k)
