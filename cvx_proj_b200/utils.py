"""Global-homography warp of the reference's README pipeline.

Mirror of ``image_warping`` in the reference's ``pyviz/utils.py:93-127`` -- same name, arguments and result:
the canvas is sized from the projected corners of the image to warp (``:99-112``), the image is warped onto it
with ``cv.warpPerspective`` semantics (bilinear, constant border 0, ``:114``) and the base image is either pasted
over the result (``direct_blend=True``, ``:124-125``) or mean-blended where the warp left something (``:115-123``,
the reference's "much slower" Python loop).  Warp and blend are one kernel (``k_warp_global``,
``csrc/warp_global.cu``), bit-exact with OpenCV's 8-bit fixed-point bilinear warp -- the INTER_LINEAR path with
coordinates in 1/32 pixel and 15-bit weights (``INTER_BITS = 5``, ``INTER_REMAP_COEF_BITS = 15``) that
``cv.warpPerspective`` takes for ``CV_8UC3``; the golden vectors (``tests/golden/ref_image_warping.npz``) were
produced with opencv-python 4.13.0, and OpenCV 3.x / 4.x releases share that path (a build that routes 8-bit images
through a different bilinear kernel would differ by +-1 in some pixels).  The O(1) host arithmetic
(corner projection, 3x3 inverse) restates ``cv.perspectiveTransform`` / ``cv::invert`` in numpy float64.
No CPU fallback.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _runtime as rt

__all__ = ["image_warping", "warp_perspective", "warping_canvas", "invert3x3", "match_descriptors", "coarse_matching"]


def invert3x3(m) -> np.ndarray:
    """Closed-form 3x3 inverse in the operation order of ``cv::invert`` (what ``cv.warpPerspective`` applies to
    its matrix); a singular matrix gives zeros like OpenCV."""
    m = np.asarray(m, dtype=np.float64)
    (a00, a01, a02), (a10, a11, a12), (a20, a21, a22) = m
    d = a00 * (a11 * a22 - a12 * a21) - a01 * (a10 * a22 - a12 * a20) + a02 * (a10 * a21 - a11 * a20)
    if d == 0:
        return np.zeros((3, 3))
    d = 1.0 / d
    return np.array([[(a11 * a22 - a12 * a21) * d, (a02 * a21 - a01 * a22) * d, (a01 * a12 - a02 * a11) * d],
                     [(a12 * a20 - a10 * a22) * d, (a00 * a22 - a02 * a20) * d, (a02 * a10 - a00 * a12) * d],
                     [(a10 * a21 - a11 * a20) * d, (a01 * a20 - a00 * a21) * d, (a00 * a11 - a01 * a10) * d]])


def warping_canvas(base_shape, warp_shape, H):
    """``(canvas_w, canvas_h, t_x, t_y, Ht.dot(H))`` of pyviz/utils.py:99-112: the four corners of the image to
    warp go through ``H`` (``cv.perspectiveTransform``: float64 inside, float32 out), the canvas is their bounding
    box united with the base image, ``np.int32(min - 0.5)`` / ``np.int32(max + 0.5)``."""
    h1, w1 = base_shape[:2]
    h2, w2 = warp_shape[:2]
    m = np.asarray(H, dtype=np.float64)
    pts2 = np.array([[0, 0], [0, h2], [w2, h2], [w2, 0]], dtype=np.float32).astype(np.float64)
    z = m[2, 0] * pts2[:, 0] + m[2, 1] * pts2[:, 1] + m[2, 2]
    z = np.where(z != 0, 1.0 / np.where(z != 0, z, 1.0), 0.0)
    pts2_ = np.stack([(m[0, 0] * pts2[:, 0] + m[0, 1] * pts2[:, 1] + m[0, 2]) * z,
                      (m[1, 0] * pts2[:, 0] + m[1, 1] * pts2[:, 1] + m[1, 2]) * z], axis=1).astype(np.float32)
    pts = np.concatenate([np.array([[0, 0], [0, h1], [w1, h1], [w1, 0]], dtype=np.float32), pts2_], axis=0)
    xmin, ymin = np.int32(pts.min(axis=0) - 0.5)
    xmax, ymax = np.int32(pts.max(axis=0) + 0.5)
    tx, ty = int(-xmin), int(-ymin)
    ht = np.array([[1, 0, tx], [0, 1, ty], [0, 0, 1]])
    return int(xmax - xmin), int(ymax - ymin), tx, ty, ht.dot(np.asarray(H))


def _device_image(torch, device, img):
    if isinstance(img, np.ndarray):
        if img.ndim != 3 or img.shape[2] != 3:
            raise ValueError("expected an [H, W, 3] image")
        return rt.to_device(torch, device, img.astype(np.uint8, copy=False))
    return img.contiguous()


def warp_perspective(src, M, dsize, base=None, offset=(0, 0), mode=0, device=None, out=None):
    """``cv.warpPerspective(src, M, dsize)`` for ``uint8 [h, w, 3]`` images (bilinear, constant border 0), optionally
    fused with the paste (``mode=1``) or mean blend (``mode=2``) of ``base`` at ``offset`` on the canvas.  numpy in ->
    numpy out; CUDA tensors in -> the canvas stays on the device."""
    on_device = not isinstance(src, np.ndarray)
    torch, device = rt.torch_cuda(src.device if on_device else device)
    lib = rt.load_library()
    width, height = int(dsize[0]), int(dsize[1])
    src_dev = _device_image(torch, device, src)
    base_dev = _device_image(torch, device, base) if base is not None else None
    if out is None:
        out = torch.empty((height, width, 3), dtype=torch.uint8, device=device)
    minv = np.ascontiguousarray(invert3x3(M), dtype=np.float64)
    bh, bw = (base_dev.shape[0], base_dev.shape[1]) if base_dev is not None else (0, 0)
    with torch.cuda.device(device):
        rt.check(lib.apap_warp_perspective(
            src_dev.data_ptr(), src_dev.shape[0], src_dev.shape[1], minv.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
            out.data_ptr(), height, width, base_dev.data_ptr() if base_dev is not None else None, bh, bw,
            int(offset[0]), int(offset[1]), int(mode), rt.stream_ptr(torch, device)), "apap_warp_perspective")
    return out if on_device else rt.to_host(torch, out)


def image_warping(img_base, img2warp, H, direct_blend=True, device=None):
    """Warp ``img2warp`` by the homography ``H`` onto a canvas that also holds ``img_base`` (pyviz/utils.py:93-127).

    ``direct_blend=True``: the base image covers the warped one; ``False``: the mean of the two where the warp is
    non-empty (any channel > 0), the base pixel elsewhere -- the ghosting view of the reference."""
    cw, ch, tx, ty, m = warping_canvas(img_base.shape, img2warp.shape, H)
    return warp_perspective(img2warp, m, (cw, ch), base=img_base, offset=(tx, ty), mode=1 if direct_blend else 2,
                            device=device)


# ------------------------------------------------------------------ keypoint-pair producer
def match_descriptors(feats_query, feats_train, device=None):
    """Exact 1-nearest-neighbour match of every query descriptor in the train set on the GPU (``apap_match_nn``):
    ``(train_idx int32 [nq], distance float32 [nq])`` = what ``cv.BFMatcher(cv.NORM_L2).match(feats_query, feats_train)``
    returns as ``DMatch.trainIdx`` / ``.distance`` -- bit for bit for integer-valued descriptors such as SIFT's.
    The reference calls ``cv.FlannBasedMatcher().match`` (pyviz/utils.py:149-150), an approximate search whose result
    changes from run to run; the exact neighbour is what it approximates.  Arrays or CUDA tensors ``[n, dim]`` float32,
    ``dim`` even and <= 256."""
    torch, device = rt.torch_cuda(device if device is not None else (feats_query.device if hasattr(feats_query, "device") and
                                                                    not isinstance(feats_query, np.ndarray) else None))
    lib = rt.load_library()

    def dev(a):
        if isinstance(a, np.ndarray):
            return rt.to_device(torch, device, np.ascontiguousarray(a, dtype=np.float32))
        if a.dtype != torch.float32:
            raise TypeError("device descriptors must be float32 tensors")
        return a.contiguous()
    host_out = isinstance(feats_query, np.ndarray)
    q, t = dev(feats_query), dev(feats_train)
    if q.dim() != 2 or t.dim() != 2 or (t.shape[0] and q.shape[1] != t.shape[1]):
        raise ValueError("match_descriptors: descriptors are [n, dim] arrays of one dim")
    nq, nt, dim = int(q.shape[0]), int(t.shape[0]), int(q.shape[1])
    idx = torch.empty(nq, dtype=torch.int32, device=device)
    dist = torch.empty(nq, dtype=torch.float32, device=device)
    scratch = torch.empty(max(nq, 1), dtype=torch.int64, device=device)
    with torch.cuda.device(device):
        rt.check(lib.apap_match_nn(q.data_ptr(), t.data_ptr() if nt else None, nq, nt, dim, scratch.data_ptr(),
                                   idx.data_ptr(), dist.data_ptr(), rt.stream_ptr(torch, device)), "apap_match_nn")
    if host_out:
        return rt.to_host(torch, idx), rt.to_host(torch, dist)
    return idx, dist


def coarse_matching(c_img, o_img, raw_kpts_cp, raw_kpts_op, device=None):
    """``coarse_matching`` of the reference (pyviz/utils.py:142-151) -- same arguments, same return tuple
    ``(kpts_cp, feats_cp, kpts_op, feats_op, matches)`` with ``matches`` a list of ``cv.DMatch`` (one per centre
    keypoint, in query order).  SIFT descriptors at the given keypoints are OpenCV's own (``cv.SIFT.compute``, host:
    un-vendored, no bit-level restatement exists); the matcher is the exact nearest neighbour on the GPU
    (``match_descriptors``) where the reference asks FLANN's randomised kd-trees for an approximation of it."""
    import cv2 as cv

    kpts_cp = [cv.KeyPoint(*pt, 1) for pt in raw_kpts_cp]
    kpts_op = [cv.KeyPoint(*pt, 1) for pt in raw_kpts_op]
    extractor = cv.SIFT.create(nfeatures=128)
    kpts_cp, feats_cp = extractor.compute(c_img, kpts_cp)
    kpts_op, feats_op = extractor.compute(o_img, kpts_op)
    idx, dist = match_descriptors(feats_cp, feats_op, device=device)
    matches = [cv.DMatch(i, int(j), 0, float(d)) for i, (j, d) in enumerate(zip(idx, dist)) if j >= 0]
    return kpts_cp, feats_cp, kpts_op, feats_op, matches
