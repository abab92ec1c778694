"""Spectral match weighting of the reference's README pipeline.

Mirror of ``calculate_M`` (and its helpers ``recompute_matching`` / ``match_RANSAC``) in the reference's
``pyviz/spectral_method.py:26-136`` -- same names, arguments and results.  The N x N affinity matrix over the coarse
matches (``:96-121``) is built on the GPU in the reference's float32 / float64 arithmetic and its leading singular
vector (``np.linalg.svd`` + ``|U[:, 0]|``, ``:122-123``, O(N^3) on the CPU) comes from a float64 power iteration
on the device (``csrc/spectral.cu``; the matrix is symmetric, non-negative with a positive diagonal, so that vector
is its Perron vector).  O(N D) host pieces -- the gather of OpenCV ``KeyPoint`` / ``DMatch`` lists
(``pyviz/utils.py:131-139``), descriptor normalisation, the epipolar term, the re-matching mask -- stay in numpy.
No CPU fallback for the matrix and the iteration.
"""
from __future__ import annotations

import numpy as np

from . import _runtime as rt

__all__ = ["calculate_M", "recompute_matching", "match_RANSAC", "cv_to_array", "spectral_segment_device",
           "affinity_diagonal"]

POWER_TOL = 1e-14          # max |x_new - x_old| at which the iteration stops
POWER_MAX_ITER = 20000
POWER_CHECK_EVERY = 16     # device -> host reads of the convergence measure


def cv_to_array(source, target, matches, is_pts=True):
    """pyviz/utils.py:131-139: gather ``source[m.queryIdx]`` / ``target[m.trainIdx]`` per match -- keypoint
    coordinates as float32 ``[N, 2]`` or descriptor rows stacked (copies, like the reference's ``np.stack``)."""
    if is_pts:
        src = np.float32([source[m.queryIdx].pt for m in matches])
        dst = np.float32([target[m.trainIdx].pt for m in matches])
    else:
        src = np.stack([source[m.queryIdx] for m in matches], axis=0)
        dst = np.stack([target[m.trainIdx] for m in matches], axis=0)
    return src, dst


def match_RANSAC(kpts_cp, kpts_op, matches, swap=False):
    """pyviz/spectral_method.py:26-33: ``cv.findHomography(..., cv.RANSAC, 5.0)`` on the coarse matches -- the
    reference's own OpenCV call (the keypoint-pair producer, SURVEY row N3, is not rebuilt here)."""
    import cv2 as cv

    if swap:
        dst_pts, src_pts = cv_to_array(kpts_cp, kpts_op, matches)
    else:
        src_pts, dst_pts = cv_to_array(kpts_cp, kpts_op, matches)
    H, mask = cv.findHomography(src_pts, dst_pts, cv.RANSAC, 5.0)
    return H, mask.astype(np.float32).ravel()


def recompute_matching(kpts_cp, feats_cp, kpts_op, feats_op, matches, H, opts):
    """pyviz/spectral_method.py:34-64, vectorised: 1.0 for a match whose other-image keypoint lands within
    ``opts.em_radius`` of the centre keypoint after ``H`` and whose normalised descriptors' dot product exceeds
    ``opts.score_thresh``."""
    c_pts, o_pts = cv_to_array(kpts_cp, kpts_op, matches)
    c_feats, o_feats = cv_to_array(feats_cp, feats_op, matches, is_pts=False)
    c_feats = c_feats / np.linalg.norm(c_feats, axis=-1, keepdims=True)
    o_feats = o_feats / np.linalg.norm(o_feats, axis=-1, keepdims=True)
    homo = np.concatenate([o_pts, np.ones((o_pts.shape[0], 1), dtype=np.float32)], axis=1)      # float32, like np.float32((*pt, 1))
    warped = homo @ np.asarray(H).T                                                               # float64 (H) x float32
    warped = warped[:, :2] / warped[:, 2:3]
    dist = np.linalg.norm(warped - c_pts, axis=-1)
    score = np.sum(c_feats * o_feats, axis=-1)
    return ((dist < opts.em_radius) & (score > opts.score_thresh)).astype(np.float32)


def affinity_diagonal(src_pts, dst_pts, c_feats, o_feats, F, epi_weight):
    """Diagonal of M (pyviz/spectral_method.py:100-111): descriptor similarity + ``epi_weight / (1 + |x'^T F x|)``,
    float64 ``[N]``.  O(N D) host arithmetic, written as the reference writes it."""
    n = src_pts.shape[0]
    homo_src = np.hstack((src_pts, np.ones((n, 1))))
    homo_dst = np.hstack((dst_pts, np.ones((n, 1))))
    c_feats = c_feats / np.linalg.norm(c_feats, axis=-1, keepdims=True)
    o_feats = o_feats / np.linalg.norm(o_feats, axis=-1, keepdims=True)
    epi_vectors = np.asarray(F) @ homo_src.T
    epi_score = np.abs(np.sum(homo_dst * epi_vectors.T, axis=-1))
    match_score = np.sum(c_feats * o_feats, axis=-1)
    return np.asarray(match_score + epi_weight / (1.0 + epi_score), dtype=np.float64)


def spectral_segment_device(src_pts, dst_pts, diag, affinity_eps, device=None, return_info=False):
    """``|U[:, 0]| / max`` of the affinity matrix (pyviz/spectral_method.py:112-125) on the GPU: matrix build +
    power iteration.  ``src_pts`` / ``dst_pts``: float32 ``[N, 2]``; ``diag``: float64 ``[N]``.  N <= 65535 and the dense
    float64 matrix must fit the device (``ApapError`` otherwise); a ``RuntimeWarning`` (and ``info["converged"] = False``)
    when the iteration has not settled in ``POWER_MAX_ITER`` steps (repeated leading eigenvalue)."""
    torch, device = rt.torch_cuda(device)
    lib = rt.load_library()
    n = int(src_pts.shape[0])
    if n == 0:
        return (np.zeros(0), {"iterations": 0}) if return_info else np.zeros(0)
    if n > 65535:
        raise rt.ApapError(f"spectral_segment_device: {n} matches; the dense affinity matrix kernels take at most 65535")
    free_bytes, _ = torch.cuda.mem_get_info(device)
    if 8 * n * n > 0.9 * free_bytes:
        raise rt.ApapError(f"spectral_segment_device: the {n} x {n} float64 affinity matrix needs {8 * n * n / 2**30:.1f} GiB, "
                           f"{free_bytes / 2**30:.1f} GiB are free on {device}")
    rcp_value = np.float32(1 / 2 / (affinity_eps ** 2))          # the float32 the reference's float32 product sees
    s_dev = rt.to_device(torch, device, np.ascontiguousarray(src_pts, dtype=np.float32))
    d_dev = rt.to_device(torch, device, np.ascontiguousarray(dst_pts, dtype=np.float32))
    g_dev = rt.to_device(torch, device, np.ascontiguousarray(diag, dtype=np.float64))
    m = torch.empty((n, n), dtype=torch.float64, device=device)
    y = torch.empty((2, n), dtype=torch.float64, device=device)              # ping-pong iterates
    y[0].fill_(1.0 / np.sqrt(n))
    norms = torch.tensor([1.0, 0.0, 0.0], dtype=torch.float64, device=device)   # ring of |y_k|^2
    x = torch.empty(n, dtype=torch.float64, device=device)
    diff = torch.zeros(1, dtype=torch.int64, device=device)     # bits of a non-negative double
    iters, converged, last = 0, True, float("inf")
    with torch.cuda.device(device):
        st = rt.stream_ptr(torch, device)
        rt.check(lib.apap_affinity_matrix(s_dev.data_ptr(), d_dev.data_ptr(), g_dev.data_ptr(), n, float(rcp_value),
                                          m.data_ptr(), st), "apap_affinity_matrix")
        while iters < POWER_MAX_ITER:
            rt.check(lib.apap_power_iterate(m.data_ptr(), n, y.data_ptr(), norms.data_ptr(), iters, POWER_CHECK_EVERY,
                                            x.data_ptr(), diff.data_ptr(), st), "apap_power_iterate")
            iters += POWER_CHECK_EVERY
            last = diff.view(torch.float64).item()
            if last <= POWER_TOL:
                break
        else:
            # two (nearly) equal leading eigenvalues -- e.g. two disjoint, equally consistent clusters of matches: the
            # iterate is a mixture of their vectors, where np.linalg.svd returns one of them.  Say so instead of
            # handing it over silently (info["converged"] carries the same for return_info callers).
            import warnings
            warnings.warn(f"spectral_segment_device: power iteration stopped at {iters} steps with max|dx| = {last:.2e} "
                          f"(> {POWER_TOL:g}); the leading eigenvalue of the affinity matrix is (nearly) repeated and the "
                          "scores are a mixture of its eigenvectors", RuntimeWarning, stacklevel=2)
            converged = False
    seg = np.abs(rt.to_host(torch, x))
    seg /= np.max(seg)
    seg[seg < 1e-6] = 0
    return (seg, {"iterations": iters, "converged": converged, "last_step": last}) if return_info else seg


def calculate_M(kpts_cp, feats_cp, kpts_op, feats_op, F, matches, opts, verbose=False, swap=True, init_ransac=True,
                Hg=None, device=None):
    """Per-match spectral score and the mask it refines (pyviz/spectral_method.py:66-136).

    Returns ``(segment, H, ransac_mask, original_mask)`` like the reference: ``segment`` float64 ``[N]`` in [0, 1];
    ``H`` the homography used for the initial mask (``Hg``, or RANSAC's); ``ransac_mask`` float32 = ``aff_thresh`` x
    the initial mask with ``segment`` written where it exceeds ``opts.aff_thresh``; ``original_mask`` the initial mask.
    ``verbose`` is accepted and ignored (the reference prints and plots M)."""
    if init_ransac:
        if Hg is not None:
            H = Hg
            ransac_mask = recompute_matching(kpts_cp, feats_cp, kpts_op, feats_op, matches, Hg, opts)
        else:
            H, ransac_mask = match_RANSAC(kpts_cp, kpts_op, matches, swap)
        original_mask = ransac_mask.copy()
    else:
        H, ransac_mask, original_mask = None, None, None
    src_pts, dst_pts = cv_to_array(kpts_cp, kpts_op, matches)
    c_feats, o_feats = cv_to_array(feats_cp, feats_op, matches, is_pts=False)
    diag = affinity_diagonal(src_pts, dst_pts, c_feats, o_feats, F, opts.epi_weight)
    segment = spectral_segment_device(src_pts, dst_pts, diag, opts.affinity_eps, device=device)
    bool_mask = segment > opts.aff_thresh
    ransac_mask *= opts.aff_thresh            # TypeError with init_ransac=False, like the reference (:135)
    ransac_mask[bool_mask] = segment[bool_mask]
    return segment, H, ransac_mask, original_mask
