"""CPU oracle for the spectral match weighting -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

SURVEY.md section 8f row N4 (second half): ``calculate_M`` of the reference's ``pyviz/spectral_method.py:66-133``
-- the N x N affinity matrix over the coarse matches (diagonal: descriptor similarity + epipolar score, off
diagonal: pairwise distance consistency, ``:96-121``), its leading singular vector as a per-match score
(``:122-125``) and the replacement of the RANSAC / re-matching mask by that score (``:134-136``), with
``recompute_matching`` (``:34-64``) as the mask producer when a global homography is given.

Array-level restatement: OpenCV ``KeyPoint`` / ``DMatch`` lists are already gathered into arrays
(``pyviz/utils.py:131-139`` ``cv_to_array``).  Parity pinning: ``tests/golden/ref_spectral.npz`` holds outputs of the
LIVE reference function (``oracle/gen_golden_spectral.py`` imports ``/root/reference/pyviz/spectral_method.py``
unmodified, with empty stand-ins for the modules it imports but ``calculate_M`` never touches -- matplotlib,
configargparse, cvxpy -- which are not installed here).  Only ``tests/`` may import this module.
"""
from __future__ import annotations

import numpy as np


def recompute_matching(c_pts, o_pts, c_feats, o_feats, hmat, em_radius, score_thresh):
    """pyviz/spectral_method.py:34-64 for gathered arrays: 1 where the other image's keypoint lands within
    ``em_radius`` of the centre keypoint after ``hmat`` AND the normalised descriptors' dot product exceeds
    ``score_thresh``.  ``c_pts`` / ``o_pts``: ``[N, 2]`` float32; feats ``[N, D]`` float32."""
    n = c_pts.shape[0]
    mask = np.zeros(n, dtype=np.float32)
    for i in range(n):
        feat_c = c_feats[i] / np.linalg.norm(c_feats[i])
        feat_o = o_feats[i] / np.linalg.norm(o_feats[i])
        kpt_c = np.float32(c_pts[i])
        kpt_o = np.matmul(hmat, np.float32((o_pts[i, 0], o_pts[i, 1], 1)))
        kpt_o = (kpt_o / kpt_o[2])[:-1]
        dist = np.linalg.norm(kpt_o - kpt_c)
        feat_score = np.sum(feat_c * feat_o)
        if dist < em_radius and feat_score > score_thresh:
            mask[i] = 1.0
    return mask


def affinity_matrix(src_pts, dst_pts, c_feats, o_feats, fmat, epi_weight, affinity_eps):
    """pyviz/spectral_method.py:96-121 -- ``M [N, N]`` float64 (off-diagonal entries carry float32 arithmetic)."""
    n = src_pts.shape[0]
    homo_src = np.hstack((src_pts, np.ones((n, 1))))
    homo_dst = np.hstack((dst_pts, np.ones((n, 1))))
    c = c_feats / np.linalg.norm(c_feats, axis=-1, keepdims=True)
    o = o_feats / np.linalg.norm(o_feats, axis=-1, keepdims=True)
    epi_vectors = fmat @ homo_src.T
    epi_score = np.abs(np.sum(homo_dst * epi_vectors.T, axis=-1))
    match_score = np.sum(c * o, axis=-1)
    m = np.diag(match_score + epi_weight / (1.0 + epi_score))
    rcp_value = 1 / 2 / (affinity_eps ** 2)
    src_matrix = np.sum((src_pts.reshape(-1, 1, 2) - src_pts.reshape(1, -1, 2)) ** 2, axis=-1)
    dst_matrix = np.sum((dst_pts.reshape(-1, 1, 2) - dst_pts.reshape(1, -1, 2)) ** 2, axis=-1)
    dist_matrix = ((src_matrix - dst_matrix) ** 2) * rcp_value
    off_score = np.maximum(4.5 - dist_matrix, 0.0)
    np.fill_diagonal(off_score, 0.0)
    m += off_score
    return m


def spectral_segment(m):
    """pyviz/spectral_method.py:122-125 -- ``|U[:, 0]|`` of the SVD, scaled to max 1, entries below 1e-6 zeroed."""
    u, _, _ = np.linalg.svd(m)
    segment = np.abs(u[:, 0])
    segment /= np.max(segment)
    segment[segment < 1e-6] = 0
    return segment


def calculate_m(src_pts, dst_pts, c_feats, o_feats, fmat, hg, *, epi_weight, affinity_eps, aff_thresh, em_radius,
                score_thresh):
    """``calculate_M(..., init_ransac=True, Hg=hg)`` -> ``(segment, ransac_mask, original_mask)``
    (pyviz/spectral_method.py:87-136; ``src_pts`` = centre keypoints, ``dst_pts`` = the other image's)."""
    ransac_mask = recompute_matching(src_pts, dst_pts, c_feats, o_feats, hg, em_radius, score_thresh)
    original_mask = ransac_mask.copy()
    segment = spectral_segment(affinity_matrix(src_pts, dst_pts, c_feats, o_feats, fmat, epi_weight, affinity_eps))
    bool_mask = segment > aff_thresh
    ransac_mask *= aff_thresh
    ransac_mask[bool_mask] = segment[bool_mask]
    return segment, ransac_mask, original_mask
