"""Golden vectors of the global-homography warp from the LIVE reference.

    python oracle/gen_golden_warping.py        (build container only: needs /root/reference and cv2)

Calls ``image_warping`` of ``/root/reference/pyviz/utils.py`` (unmodified; ``:93-127``) on seeded synthetic
images and homographies and stores its outputs -- whole canvases for the small cases, row CRCs + SHA-256 for
the larger ones -- in ``tests/golden/ref_image_warping.npz``.  The inputs are regenerated from the seeds by
``warping_case`` below, which the tests import.
"""
from __future__ import annotations

import hashlib
import os
import sys
import zlib

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from cvx_proj_b200 import synth  # noqa: E402

# name -> (width, height, extra width / height of the image to warp, strength of the perturbation of H)
CASES = {
    "w64": (64, 48, 0, 0, 1.0),
    "w160": (160, 120, 2, 1, 2.0),
    "thin": (33, 7, 1, 0, 1.0),
    "tall": (5, 300, 0, 1, 3.0),
    "odd": (257, 191, 1, 1, 2.0),
    "vga": (640, 480, 0, 0, 1.0),
    "c1": (1024, 768, 0, 0, 1.0),
}
FULL = ("w64", "w160", "thin", "tall")        # canvases stored whole


def warping_case(name):
    """``(img_base, img2warp, H)`` of a named case (seeded)."""
    w, h, ew, eh, k = CASES[name]
    seed = sorted(CASES).index(name)
    rng = np.random.default_rng(1000 + seed)
    hmat = synth.ground_truth_h(w, h).copy()
    hmat[:2, :2] += rng.normal(0, 0.05, (2, 2)) * k
    hmat[:2, 2] += rng.normal(0, 30, 2) * k * (w / 640.0)
    hmat[2, :2] += rng.normal(0, 2e-5, 2) * k * (640.0 / w)
    return synth.make_image(w, h, seed=seed), synth.make_image(w + ew, h + eh, seed=100 + seed), hmat


def main():
    sys.path.insert(0, "/root/reference/pyviz")
    np.int = int  # noqa: alias removed from numpy, used elsewhere in the reference's modules
    import cv2
    import utils as ref_utils  # the reference, unmodified

    out = {"versions": np.array([f"numpy {np.__version__}", f"opencv {cv2.__version__}"])}
    for name in CASES:
        base, warp, hmat = warping_case(name)
        for db in (True, False):
            res = ref_utils.image_warping(base, warp, hmat, direct_blend=db)
            tag = f"{name}_{'paste' if db else 'mean'}"
            out[tag + "_shape"] = np.array(res.shape)
            out[tag + "_sha"] = np.array(hashlib.sha256(res.tobytes()).hexdigest())
            out[tag + "_rowcrc"] = np.array([zlib.crc32(np.ascontiguousarray(r).tobytes()) for r in res], dtype=np.uint32)
            if name in FULL:
                out[tag] = res
        out[name + "_H"] = hmat
    path = os.path.join(REPO, "tests", "golden", "ref_image_warping.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
