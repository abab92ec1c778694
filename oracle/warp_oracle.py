"""CPU oracle for the global-homography warp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

SURVEY.md section 8f row N4: ``image_warping`` of the reference's ``pyviz/utils.py:93-127`` -- the canvas
from the projected corners, ``cv.warpPerspective`` (bilinear, constant border 0) of the image to warp, then
either the base image pasted over it or the mean blend of the reference's Python loop.

``cv.warpPerspective`` is un-vendored third-party arithmetic (opencv-python >= 4.5 in the reference's
``requirements.txt``; 4.13.0 in the build container).  What is restated here is OpenCV's published
fixed-point algorithm for 8-bit bilinear warps (``modules/imgproc/src/imgwarp.cpp``,
``WarpPerspectiveInvoker`` + ``remapBilinear``):
  * the matrix is inverted with the closed-form 3x3 inverse of ``cv::invert``;
  * destination pixels are processed in blocks ``bw0`` wide; for a pixel ``x = xb + x1`` of row ``y``
    ``X0 = M0 xb + M1 y + M2`` (float64), ``W = 32 / (W0 + M6 x1)``, ``X = cvRound(clamp((X0 + M0 x1) W))`` --
    source coordinates in 1/32 pixel;
  * the four taps ``(X >> 5, Y >> 5) + {0, 1}^2`` are weighted by ``(1 - fx)(1 - fy) ...`` with ``fx = (X & 31) / 32``
    in 15-bit fixed point (exact: the products are multiples of 2^-10) and the sum is rounded,
    ``(sum + 2^14) >> 15``; a tap outside the source contributes the border value 0.
Parity pinning: ``tests/golden/ref_image_warping.npz`` holds outputs of the LIVE reference function
(``oracle/gen_golden_warping.py`` imports ``/root/reference/pyviz/utils.py`` unmodified); this module reproduces
them bit for bit (``tests/test_oracle_golden.py``).

Only ``tests/``, ``smoke()`` and ``bench.py``'s CPU legs may import this module.
"""
from __future__ import annotations

import numpy as np

INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS
_INT_MIN, _INT_MAX = -2147483648.0, 2147483647.0


def invert3x3(m):
    """``cv::invert`` of a 3x3 float64 matrix (closed form, the operation order of ``core/src/lapack.cpp``)."""
    m = np.asarray(m, dtype=np.float64)
    (a00, a01, a02), (a10, a11, a12), (a20, a21, a22) = m
    d = a00 * (a11 * a22 - a12 * a21) - a01 * (a10 * a22 - a12 * a20) + a02 * (a10 * a21 - a11 * a20)
    if d == 0:
        return np.zeros((3, 3))
    d = 1.0 / d
    return np.array([[(a11 * a22 - a12 * a21) * d, (a02 * a21 - a01 * a22) * d, (a01 * a12 - a02 * a11) * d],
                     [(a12 * a20 - a10 * a22) * d, (a00 * a22 - a02 * a20) * d, (a02 * a10 - a00 * a12) * d],
                     [(a10 * a21 - a11 * a20) * d, (a01 * a20 - a00 * a21) * d, (a00 * a11 - a01 * a10) * d]])


def block_width(width: int, height: int) -> int:
    """Width of the destination blocks of ``WarpPerspectiveInvoker`` (BLOCK_SZ = 32)."""
    bh0 = min(16, height)
    return min(1024 // max(bh0, 1), width)


def fixed_point_coords(minv, width: int, height: int):
    """``(X, Y)`` int64 ``[height, width]``: source coordinates in 1/32 pixel, as OpenCV computes them."""
    m = np.asarray(minv, dtype=np.float64).reshape(9)
    bw0 = max(block_width(width, height), 1)
    x = np.arange(width, dtype=np.int64)
    xb = ((x // bw0) * bw0).astype(np.float64)[None, :]
    x1 = (x % bw0).astype(np.float64)[None, :]
    y = np.arange(height, dtype=np.float64)[:, None]
    x0 = (m[0] * xb + m[1] * y) + m[2]
    y0 = (m[3] * xb + m[4] * y) + m[5]
    w0 = (m[6] * xb + m[7] * y) + m[8]
    w = w0 + m[6] * x1
    with np.errstate(divide="ignore", invalid="ignore"):
        w = np.where(w != 0, INTER_TAB_SIZE / w, 0.0)
        fx = np.maximum(_INT_MIN, np.minimum(_INT_MAX, (x0 + m[0] * x1) * w))
        fy = np.maximum(_INT_MIN, np.minimum(_INT_MAX, (y0 + m[3] * x1) * w))
    fx = np.where(np.isnan(fx), _INT_MIN, fx)          # cvRound of NaN = INT_MIN (the x86 "indefinite" integer)
    fy = np.where(np.isnan(fy), _INT_MIN, fy)
    return np.rint(fx).astype(np.int64), np.rint(fy).astype(np.int64)


def warp_perspective(src, m, dsize):
    """``cv.warpPerspective(src, m, dsize)`` for ``uint8 [h, w, 3]``: bilinear, BORDER_CONSTANT 0."""
    width, height = int(dsize[0]), int(dsize[1])
    src = np.asarray(src)
    sh, sw = src.shape[:2]
    xq, yq = fixed_point_coords(invert3x3(m), width, height)
    sx = np.clip(xq >> INTER_BITS, -32768, 32767)
    sy = np.clip(yq >> INTER_BITS, -32768, 32767)
    ax = (xq & (INTER_TAB_SIZE - 1)).astype(np.int64)
    ay = (yq & (INTER_TAB_SIZE - 1)).astype(np.int64)
    acc = np.zeros((height, width, src.shape[2]), dtype=np.int64)
    for dy in (0, 1):
        for dx in (0, 1):
            wgt = (ax if dx else INTER_TAB_SIZE - ax) * (ay if dy else INTER_TAB_SIZE - ay) * 32   # x 2^15 / 2^10
            xx, yy = sx + dx, sy + dy
            ok = (xx >= 0) & (xx < sw) & (yy >= 0) & (yy < sh)
            tap = np.zeros_like(acc)
            tap[ok] = src[yy[ok], xx[ok]]
            acc += wgt[..., None] * tap
    return ((acc + (1 << 14)) >> 15).astype(np.uint8)


def perspective_transform_corners(w: int, h: int, hmat):
    """``cv.perspectiveTransform`` of the four corners ``(0,0) (0,h) (w,h) (w,0)`` as float32 (float64 inside)."""
    pts = np.array([[0, 0], [0, h], [w, h], [w, 0]], dtype=np.float32).astype(np.float64)
    m = np.asarray(hmat, dtype=np.float64)
    x = m[0, 0] * pts[:, 0] + m[0, 1] * pts[:, 1] + m[0, 2]
    y = m[1, 0] * pts[:, 0] + m[1, 1] * pts[:, 1] + m[1, 2]
    z = m[2, 0] * pts[:, 0] + m[2, 1] * pts[:, 1] + m[2, 2]
    z = np.where(z != 0, 1.0 / z, 0.0)
    return np.stack([x * z, y * z], axis=1).astype(np.float32)


def warping_canvas(base_shape, warp_shape, hmat):
    """pyviz/utils.py:99-112 -- ``(canvas_w, canvas_h, t_x, t_y, Ht . H)``."""
    h1, w1 = base_shape[:2]
    h2, w2 = warp_shape[:2]
    pts1 = np.array([[0, 0], [0, h1], [w1, h1], [w1, 0]], dtype=np.float32)
    pts = np.concatenate([pts1, perspective_transform_corners(w2, h2, hmat)], axis=0)
    xmin, ymin = np.int32(pts.min(axis=0) - 0.5)
    xmax, ymax = np.int32(pts.max(axis=0) + 0.5)
    tx, ty = int(-xmin), int(-ymin)
    ht = np.array([[1, 0, tx], [0, 1, ty], [0, 0, 1]])
    return int(xmax - xmin), int(ymax - ymin), tx, ty, ht.dot(np.asarray(hmat))


def image_warping(img_base, img2warp, hmat, direct_blend=True):
    """pyviz/utils.py:93-127, the mean-blend loop vectorised (float32 mean then truncation = ``(a + b) >> 1``)."""
    cw, ch, tx, ty, m = warping_canvas(img_base.shape, img2warp.shape, hmat)
    result = warp_perspective(img2warp, m, (cw, ch))
    h1, w1 = img_base.shape[:2]
    region = result[ty:h1 + ty, tx:w1 + tx]
    if direct_blend:
        region[...] = img_base
    else:
        filled = region.any(axis=-1, keepdims=True)
        mean = ((img_base.astype(np.float32) + region.astype(np.float32)) / 2.0).astype(np.uint8)
        region[...] = np.where(filled, mean, img_base)
    return result
