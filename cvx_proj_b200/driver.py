"""The reference's APAP driver (``pyviz/apap.py:220-265``, its ``__main__`` block) as functions.

The reference script has no function for what it does after ``local_homography``: the per-cell
inverse + ``/[2, 2]`` + column-major ``[cells, 9]`` float64 layout of the ``.mat`` product
(``:250-265``), and -- commented out -- the warp / paste / blend that makes the stitched image
(``:258-262``).  A caller that switches packages needs both, so they live here:

  ``mat_layout``      pyviz/apap.py:250-254,263-264   (in place on the caller's grid, like the script)
  ``save2mat``        pyviz/utils.py:68-70
  ``stitch_pair``     pyviz/apap.py:238-265           (grid, H field, ``.mat`` matrix, stitched image)

The keypoint-pair producer (``visualize_feature_pairs``) and the dataset IO stay with the caller
(SURVEY.md section 8f, rows N3/N4): ``stitch_pair`` starts where the script has its matched
keypoints and its two images.
"""
from __future__ import annotations

from typing import NamedTuple, Optional

import numpy as np

from .apap_utils import final_size, get_mesh, get_vertice

__all__ = ["mat_layout", "save2mat", "stitch_pair", "StitchResult"]


def mat_layout(local_homography: np.ndarray, stitcher=None) -> np.ndarray:
    """``[mesh_y, mesh_x, 3, 3]`` float32 grid -> the ``[cells, 9]`` float64 matrix the script saves.

    Like the script (pyviz/apap.py:250-254) it first replaces every cell of ``local_homography``
    IN PLACE by its float32 inverse divided by its last entry; one stacked ``np.linalg.inv`` is the
    same LAPACK call per 3x3 block as the reference's loop (bit-identical, tests).  Then
    ``transpose(0, 1, 3, 2)`` (column-major 3x3, what the MATLAB evaluator reads), float64,
    ``reshape(-1, 9)`` (pyviz/apap.py:263-264).  With ``stitcher`` (an ``APAP``) the inverse runs on its GPU
    (``APAP.invert_grid``: same bits, 10 ms less at 40 000 cells).
    """
    if stitcher is not None:
        stitcher.invert_grid(local_homography)
    else:
        local_homography[...] = np.linalg.inv(local_homography)
    local_homography /= local_homography[..., -1:, -1:].copy()
    return local_homography.transpose(0, 1, 3, 2).astype(np.float64).reshape(-1, 9)


def save2mat(path: str, arr: np.ndarray, name: str = "sift_feature", prefix: str = "./output/") -> str:
    """``scipy.io.savemat(f"{prefix}{path}.mat", {name: arr})`` (pyviz/utils.py:68-70); returns the file name.
    The script calls it as ``save2mat(f"case{c}/H3{i}_apap", H, name='H', prefix="../diff_1/results/")``."""
    import scipy.io

    mat_file_path = f"{prefix}{path}.mat"
    scipy.io.savemat(mat_file_path, {name: arr})
    return mat_file_path


class StitchResult(NamedTuple):
    final_size: tuple            # (final_w, final_h, offset_x, offset_y)   pyviz/apap.py:238
    mesh: np.ndarray             # [2, mesh_size + 1] float64 cell edges      :239
    vertices: np.ndarray         # [mesh_size, mesh_size, 2] float64 anchors  :240
    local_homography: np.ndarray  # [mesh, mesh, 3, 3] float32 as returned by APAP.local_homography   :242
    mat: np.ndarray              # [cells, 9] float64, the matrix the script saves under 'H'        :263-265
    stitched: Optional[np.ndarray]  # [final_h, final_w, 3] uint8 blend of the warped and the centre image, or None


def stitch_pair(center_img, other_img, final_src, final_dst, project_h, *, mesh_size: int = 100, gamma=0.5,
                sigma=100, warp_img=None, blend_img=None, stitch: bool = True, device=None) -> StitchResult:
    """One pass of the script from its matched keypoints on (pyviz/apap.py:238-265).

    ``center_img`` / ``other_img`` size the canvas (``final_size``), ``final_src`` / ``final_dst`` are the
    matched keypoints (other image, centre image) and ``project_h`` the global homography, all as
    ``visualize_feature_pairs(..., swap=True)`` returns them.  ``warp_img`` / ``blend_img`` are the
    images that get warped and pasted (the script uses the de-hazed pair, ``:245``; default: the same
    two images).  With ``stitch`` the commented pipeline runs too: warp the other image through the
    per-cell homographies, paste the centre image at the offsets, ``uniform_blend`` -- one fused kernel.
    """
    from .apap import APAP

    final_w, final_h, offset_x, offset_y = final_size(center_img, other_img, project_h)
    mesh = get_mesh((final_w, final_h), mesh_size + 1)
    vertices = get_vertice((final_w, final_h), mesh_size, (offset_x, offset_y))
    stitcher = APAP(gamma, sigma, [final_w, final_h], [offset_x, offset_y], device=device)
    local_h, _ = stitcher.local_homography(final_src, final_dst, vertices)
    mat = mat_layout(local_h.copy(), stitcher)
    stitched = None
    if stitch:
        # The script's commented lines (:258-261) would hand local_warp the grid it has just inverted for the
        # .mat product, and local_warp inverts its argument again (:201-203).  Here the warp gets the grid as
        # local_homography returned it -- canvas pixel -> inv(H_cell) -> pixel of the other image -- which is
        # the call `stitcher.local_warp(other_img, local_homography, mesh)` made right after :242.
        warp_img = other_img if warp_img is None else warp_img
        blend_img = center_img if blend_img is None else blend_img
        stitched = stitcher.local_warp_blend(warp_img, local_h.copy(), mesh, blend_img)
    return StitchResult((final_w, final_h, offset_x, offset_y), mesh, vertices, local_h, mat, stitched)
