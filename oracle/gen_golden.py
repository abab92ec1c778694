"""Generate the golden vectors under ``tests/golden/`` from the LIVE reference.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python oracle/gen_golden.py

It imports ``/root/reference/pyviz/apap.py`` and ``apap_utils.py`` unmodified (the only
shim is ``np.int = int``, an alias numpy >= 1.24 removed and ``apap_utils.py:59`` still
uses), feeds them the seeded synthetic scenes of ``cvx_proj_b200.synth`` and stores the
reference's outputs.  The reference has no tests or fixtures of its own for this path
(SURVEY.md section 4), so these files are what pins the oracle and the CUDA path.

Versions the vectors were produced with are recorded inside each file.
"""
from __future__ import annotations

import contextlib
import hashlib
import io
import os
import sys
import time
import zlib

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, "/root/reference/pyviz")
np.int = int  # noqa: removed alias used by reference apap_utils.py:59

import cv2  # noqa: E402
import apap as ref_apap  # noqa: E402  (the reference, unmodified)
import apap_utils as ref_utils  # noqa: E402

from cvx_proj_b200 import synth  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def _versions():
    return np.array([f"numpy {np.__version__}", f"opencv {cv2.__version__}",
                     f"python {sys.version.split()[0]}"])


def _row_crc(img):
    return np.array([zlib.crc32(np.ascontiguousarray(r).tobytes()) for r in img], dtype=np.uint32)


def small_case(name, warp=True):
    sc = synth.make_scene(name)
    img = sc.image(1)
    centre = synth.make_image(sc.width, sc.height, seed=2)
    # host pieces, one by one
    n1, nf1 = ref_apap.APAP.getNormalize2DPts(sc.src)
    n2, nf2 = ref_apap.APAP.getNormalize2DPts(sc.dst)
    c1 = ref_apap.APAP.getConditionerFromPts(nf1)
    c2 = ref_apap.APAP.getConditionerFromPts(nf2)
    cf1 = ref_apap.APAP.point_normalize(nf1, c1)
    cf2 = ref_apap.APAP.point_normalize(nf2, c2)
    aa = ref_apap.APAP.matrix_generate(sc.src.shape[0], cf1, cf2)
    # canvas helpers
    fsz = ref_utils.final_size(img, img, sc.h_gt)
    mesh = ref_utils.get_mesh((sc.final_w, sc.final_h), sc.mesh_cells + 1)
    vert = ref_utils.get_vertice((sc.final_w, sc.final_h), sc.mesh_cells, (sc.offset_x, sc.offset_y))
    st = ref_apap.APAP(sc.gamma, sc.sigma, [sc.final_w, sc.final_h], [sc.offset_x, sc.offset_y])
    h, w = st.local_homography(sc.src, sc.dst, vert)
    out = dict(versions=_versions(), src=sc.src, dst=sc.dst, h_gt=sc.h_gt,
               final_size=np.array([int(v) for v in fsz]), mesh=mesh, vertices=vert,
               N1=n1, N2=n2, nf1=nf1, nf2=nf2, C1=c1, C2=c2, cf1=cf1, cf2=cf2, A=aa,
               H=h, W=w, gamma=sc.gamma, sigma=sc.sigma)
    # .mat product layout (pyviz/apap.py:250-265), restated inline from the driver
    g = h.copy()
    for i in range(g.shape[0]):
        for j in range(g.shape[1]):
            g[i, j] = np.linalg.inv(g[i, j].copy())
            g[i, j] /= g[i, j, -1, -1]
    out["mat"] = g.transpose(0, 1, 3, 2).astype(np.float64).reshape(-1, 9)
    if warp:
        h_mut = h.copy()
        warped = _quiet(st.local_warp, img, h_mut, mesh)
        out["H_inverted_in_place"] = h_mut              # local_warp mutates its argument
        out["warped"] = warped
        dst_temp = np.zeros_like(warped)
        dst_temp[sc.offset_y:sc.offset_y + sc.height, sc.offset_x:sc.offset_x + sc.width] = centre
        out["blended"] = ref_utils.uniform_blend(warped, dst_temp)
    np.savez_compressed(os.path.join(OUT, f"ref_{name}.npz"), **out)
    print(f"{name}: H{h.shape} canvas {sc.final_w}x{sc.final_h}")


def weight_edge_case():
    """A scene whose weights hit the gamma clamp (small sigma) and one with a far anchor."""
    sc = synth.make_scene("tiny")
    st = ref_apap.APAP(0.5, 8.0, [sc.final_w, sc.final_h], [sc.offset_x, sc.offset_y])
    h, w = st.local_homography(sc.src, sc.dst, sc.vertices)
    np.savez_compressed(os.path.join(OUT, "ref_tiny_sigma8.npz"), versions=_versions(), H=h, W=w,
                        gamma=0.5, sigma=8.0)
    print("tiny sigma=8: clamped fraction", float((w == 0.5).mean()))


def blend_edge_case():
    rng = np.random.default_rng(7)
    a = rng.integers(0, 256, size=(37, 53, 3), dtype=np.uint8)
    b = rng.integers(0, 256, size=(37, 53, 3), dtype=np.uint8)
    a[rng.random((37, 53)) < 0.3] = 0
    b[rng.random((37, 53)) < 0.3] = 0
    a[0, 0] = (0, 0, 1)            # "non-black" by a single LSB
    b[0, 0] = (255, 255, 255)
    a[0, 1] = (255, 255, 255)
    b[0, 1] = (255, 255, 255)
    np.savez_compressed(os.path.join(OUT, "ref_blend.npz"), versions=_versions(), a=a, b=b,
                        out=ref_utils.uniform_blend(a, b))
    print("blend edge case")


def c1_case():
    sc = synth.make_scene("c1")
    img = sc.image(1)
    st = ref_apap.APAP(sc.gamma, sc.sigma, [sc.final_w, sc.final_h], [sc.offset_x, sc.offset_y])
    t0 = time.time()
    h, w = st.local_homography(sc.src, sc.dst, sc.vertices)
    t_h = time.time() - t0
    h_mut = h.copy()
    t0 = time.time()
    warped = _quiet(st.local_warp, img, h_mut, sc.mesh)
    t_w = time.time() - t0
    np.savez_compressed(
        os.path.join(OUT, "ref_c1.npz"), versions=_versions(),
        src_sha=hashlib.sha256(sc.src.tobytes() + sc.dst.tobytes()).hexdigest(),
        img_sha=hashlib.sha256(img.tobytes()).hexdigest(),
        final_size=np.array([sc.final_w, sc.final_h, sc.offset_x, sc.offset_y]),
        H=h, W_sample=w[::9, ::9, ::7].copy(), H_inverted_in_place=h_mut,
        warped_row_crc=_row_crc(warped), warped_sha=hashlib.sha256(warped.tobytes()).hexdigest(),
        warped_sample=warped[::16, ::16].copy(),
        ref_seconds=np.array([t_h, t_w]))
    print(f"c1: local_homography {t_h:.2f}s ({sc.n_cells / t_h:.0f} cells/s), "
          f"local_warp {t_w:.2f}s ({sc.canvas_px / t_w / 1e6:.3f} Mpix/s)")


def big_spot_checks():
    """Reference H at a handful of cells of c2 (N=5k, 200x200): the per-cell loop of the
    reference is sliced by row (``vertices[i:i+1, cols]``), nothing in it is changed."""
    sc = synth.make_scene("c2")
    st = ref_apap.APAP(sc.gamma, sc.sigma, [sc.final_w, sc.final_h], [sc.offset_x, sc.offset_y])
    rows = np.array([0, 37, 99, 100, 163, 199])
    cols = np.array([0, 1, 50, 101, 150, 199])
    sub = sc.vertices[rows][:, cols].copy()
    t0 = time.time()
    h, _ = st.local_homography(sc.src, sc.dst, sub)
    dt = time.time() - t0
    np.savez_compressed(os.path.join(OUT, "ref_c2_spot.npz"), versions=_versions(), rows=rows, cols=cols, H=h,
                        src_sha=hashlib.sha256(sc.src.tobytes() + sc.dst.tobytes()).hexdigest(),
                        ref_seconds_per_cell=dt / h[..., 0, 0].size)
    print(f"c2 spot: {h[..., 0, 0].size} cells, {dt / h[..., 0, 0].size * 1e3:.1f} ms/cell")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    small_case("tiny")
    small_case("mini")
    weight_edge_case()
    blend_edge_case()
    c1_case()
    big_spot_checks()
