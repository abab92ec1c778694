"""TEST INFRASTRUCTURE (not shipped, never imported by cvx_proj_b200): CPU restatement of the matcher step of the
reference's keypoint-pair producer, pyviz/utils.py:149-150 (``cv.FlannBasedMatcher().match(feats_cp, feats_op)``).

The algorithm lives in OpenCV (un-vendored dependency; opencv-python 4.13.0 in this image).  FLANN's default index
(4 randomised kd-trees, 32 checks) is an APPROXIMATE nearest-neighbour search and not reproducible run to run
(tools/n3_host_cost.py), so no bit-level oracle of the reference's match list exists.  What it approximates is the
exact nearest neighbour, which OpenCV computes with ``cv.BFMatcher(cv.NORM_L2).match``: per query the train descriptor
with the smallest ``sum_k (q_k - t_k)^2`` accumulated in float32 (``normL2Sqr_``), the first one on a tie, and
``DMatch.distance = sqrt(sum)``.  ``exact_match`` restates that in numpy; the tests pin it against the live
``cv.BFMatcher`` (parity pinned for integer-valued descriptors -- SIFT's -- where every partial sum is exact).
"""
import numpy as np


def exact_match(feats_query, feats_train, chunk=256):
    """``(train_idx int32 [nq], distance float32 [nq])``; ``train_idx`` -1 for an empty train set."""
    q = np.asarray(feats_query, dtype=np.float32)
    t = np.asarray(feats_train, dtype=np.float32)
    nq = q.shape[0]
    idx = np.full(nq, -1, dtype=np.int32)
    dist = np.full(nq, np.inf, dtype=np.float32)
    if t.shape[0] == 0:
        return idx, dist
    for a in range(0, nq, chunk):
        d = q[a:a + chunk, None, :] - t[None, :, :]               # float32
        s = np.sum(d * d, axis=-1, dtype=np.float32)              # exact for integer-valued descriptors < 2^24
        j = np.argmin(s, axis=1)                                  # first minimum = lowest train index
        idx[a:a + chunk] = j
        dist[a:a + chunk] = np.sqrt(s[np.arange(s.shape[0]), j])
    return idx, dist
