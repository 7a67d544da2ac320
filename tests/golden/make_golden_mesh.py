"""Golden vectors of the surface-extraction oracle (oracle/mc_oracle.c via oracle.cpu_ref.extract_mesh): a digest of the
meshes of three seeded fields, so that neither the table nor the restated PyMCubes / Generator3D arithmetic can drift
unnoticed.  (PyMCubes itself is not installed in this image -- see DESIGN.md section 4 -- so these vectors pin the
restatement, not the binary.)   python tests/golden/make_golden_mesh.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cpu_ref  # noqa: E402


def fields():
    rng = np.random.default_rng(2024)
    ax = np.linspace(-1, 1, 12)
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing="ij")
    blob = (0.6 - np.sqrt(X ** 2 + 0.7 * Y ** 2 + 1.3 * Z ** 2) + 0.03 * rng.normal(size=X.shape)).astype(np.float32)
    noise = rng.normal(size=(7, 7, 7)).astype(np.float32)
    slab = np.where(np.abs(Z) < 0.4, 1.5, -2.0).astype(np.float32) + 0.01 * rng.normal(size=Z.shape).astype(np.float32)
    return {"blob12": blob, "noise7": noise, "slab12": slab}


def digest(v, t, keys):
    order = np.argsort(keys, kind="stable")
    vs = v[order]
    return {"nv": len(v), "nt": len(t), "v_sum": vs.sum(0), "v_abs_sum": np.abs(vs).sum(0), "v_first": vs[:4], "v_last": vs[-4:],
            "t_first": t[:6], "t_checksum": np.array([(t.astype(np.int64) * np.array([1, 3, 7])).sum()]),
            "keys_checksum": np.array([np.sort(keys).astype(np.int64).dot(np.arange(1, len(keys) + 1) % 1009)])}


if __name__ == "__main__":
    out = {}
    for name, f in fields().items():
        for thr, pad in ((0.5, 0.1), (0.3, 0.25)):
            d = digest(*cpu_ref.extract_mesh(f, thr, pad))
            for k, val in d.items():
                out[f"{name}_{thr}_{pad}_{k}"] = np.asarray(val)
    np.savez(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_mesh.npz"), **out)
    print({k: v.tolist() for k, v in out.items() if k.endswith(("_nv", "_nt"))})
