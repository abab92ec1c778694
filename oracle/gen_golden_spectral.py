"""Golden vectors of the spectral match weighting (second half) from the LIVE reference.

    python oracle/gen_golden_spectral.py       (build container only: needs /root/reference and cv2)

Imports ``/root/reference/pyviz/spectral_method.py`` unmodified and calls its ``calculate_M`` (``:66-136``) on seeded
synthetic matches.  The module imports matplotlib, configargparse and (through ``model.py``) cvxpy at load time; none
of them is installed here and ``calculate_M`` touches none of them, so empty stand-in modules are registered first.
Inputs are regenerated from the seeds by ``spectral_case`` below, which the tests import.
"""
from __future__ import annotations

import os
import sys
import types
from types import SimpleNamespace

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from cvx_proj_b200 import synth  # noqa: E402

OPTS = dict(epi_weight=0.5, affinity_eps=30.0, aff_thresh=0.5, em_radius=6.0, score_thresh=0.4)   # options.py defaults
CASES = {"s40": (40, 0.2), "s300": (300, 0.25), "s1000": (1000, 0.4), "s2500": (2500, 0.3)}


def spectral_case(name):
    """Seeded matches: ``(c_pts, o_pts, c_feats, o_feats, F, Hg)`` -- centre / other keypoints ``[N, 2]`` float32 in
    match order, SIFT-like descriptors ``[N, 128]`` float32, a fundamental-like 3x3, the global homography that maps
    the other image's keypoints onto the centre image."""
    n, outlier_share = CASES[name]
    seed = sorted(CASES).index(name)
    rng = np.random.default_rng(2000 + seed)
    o_pts, c_pts, h_gt = synth.make_keypoints(1024, 768, n, seed=50 + seed)       # other -> centre by h_gt
    c_pts = c_pts.copy()
    bad = rng.random(n) < outlier_share
    c_pts[bad] += rng.normal(0, 60, (int(bad.sum()), 2)).astype(np.float32)
    c_feats = (np.abs(rng.normal(0, 1, (n, 128))) * 50).astype(np.float32)
    o_feats = (c_feats + rng.normal(0, 8, (n, 128)).astype(np.float32)).clip(0).astype(np.float32)
    o_feats[bad] = (np.abs(rng.normal(0, 1, (int(bad.sum()), 128))) * 50).astype(np.float32)
    fmat = rng.normal(0, 1e-3, (3, 3))
    return c_pts, o_pts, c_feats, o_feats, fmat, h_gt


def main():
    for name in ("matplotlib", "matplotlib.pyplot", "configargparse", "cvxpy"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, "/root/reference/pyviz")
    np.int = int  # noqa
    import cv2
    import spectral_method as ref  # the reference, unmodified

    out = {"versions": np.array([f"numpy {np.__version__}", f"opencv {cv2.__version__}"])}
    opts = SimpleNamespace(**OPTS)
    for name in CASES:
        c_pts, o_pts, c_feats, o_feats, fmat, hg = spectral_case(name)
        n = c_pts.shape[0]
        kc = [cv2.KeyPoint(float(x), float(y), 1) for x, y in c_pts]
        ko = [cv2.KeyPoint(float(x), float(y), 1) for x, y in o_pts]
        matches = [cv2.DMatch(i, i, 0.0) for i in range(n)]
        seg, h_ret, rmask, omask = ref.calculate_M(kc, c_feats.copy(), ko, o_feats.copy(), fmat, matches, opts, swap=True,
                                                   init_ransac=True, Hg=hg)
        assert h_ret is hg
        out[name + "_segment"] = seg
        out[name + "_ransac_mask"] = rmask
        out[name + "_original_mask"] = omask
    path = os.path.join(REPO, "tests", "golden", "ref_spectral.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
