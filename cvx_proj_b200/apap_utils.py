"""Grid / canvas / blend helpers of the APAP path (host side + the blend kernel entry).

Mirror of the reference module ``pyviz/apap_utils.py`` (same names, same argument
meaning, same return layout) so that ``from apap_utils import *`` keeps working for a
caller that switches to this package.  ``get_mesh``, ``get_vertice`` and ``final_size``
are O(mesh) host arithmetic and stay in numpy float64 so their outputs are bit-identical
to the reference's; ``uniform_blend`` runs on the GPU (kernel ``k_blend`` in
``csrc/warp_blend.cu``) and has NO CPU fallback.

Reference citations (``/root/reference/``):
  get_mesh       pyviz/apap_utils.py:10-21
  get_vertice    pyviz/apap_utils.py:23-38
  final_size     pyviz/apap_utils.py:40-73
  uniform_blend  pyviz/apap_utils.py:75-88
"""
from __future__ import annotations

import numpy as np

__all__ = ["get_mesh", "get_vertice", "final_size", "uniform_blend"]


def get_mesh(size, mesh_size, start=0):
    """Cell edges of the canvas: ``[2, mesh_size]`` float64, row 0 = x edges, row 1 = y edges.

    ``size`` is ``(width, height)``.  The driver calls it with ``mesh_size = cells + 1``
    (reference pyviz/apap.py:239).  Reference: pyviz/apap_utils.py:10-21.
    """
    width, height = size
    edges = np.empty((2, int(mesh_size)), dtype=np.float64)
    edges[0] = np.linspace(start, width, mesh_size)
    edges[1] = np.linspace(start, height, mesh_size)
    return edges


def get_vertice(size, mesh_size, offsets):
    """Per-cell anchor points ``[mesh_size, mesh_size, 2]`` (x, y) float64, minus ``offsets``.

    Reproduces the reference's spacing quirk on purpose: the anchors are
    ``linspace(0, w, mesh_size) + w / (2 * mesh_size)`` (spacing ``w / (mesh_size - 1)``,
    so the last anchor lies outside the canvas).  Reference: pyviz/apap_utils.py:23-38.
    """
    width, height = size
    ax = np.linspace(0, width, mesh_size) + width / (mesh_size * 2)
    ay = np.linspace(0, height, mesh_size) + height / (mesh_size * 2)
    grid = np.empty((ay.shape[0], ax.shape[0], 2), dtype=np.float64)
    grid[..., 0] = ax[None, :]
    grid[..., 1] = ay[:, None]
    grid -= np.array(offsets)
    return grid


def final_size(src_img, dst_img, project_H):
    """Canvas extent ``(width, height, offset_x, offset_y)`` of the stitched image.

    The four corners of ``src_img`` are pushed through ``project_H`` (float32 corner
    vectors, like the reference), truncated toward zero, and unioned with the extent of
    ``dst_img``.  The reference uses the removed alias ``np.int`` for the truncation
    (pyviz/apap_utils.py:59); plain ``int`` truncation is the same operation.
    Reference: pyviz/apap_utils.py:40-73.
    """
    h, w = src_img.shape[0], src_img.shape[1]
    proj = []
    for cx, cy in ((0, 0), (0, h), (w, 0), (w, h)):
        vec = np.matmul(project_H, np.float32([cx, cy, 1]))
        proj.append([vec[0] / vec[2], vec[1] / vec[2]])
    proj = np.array(proj).astype(int)

    dh, dw = dst_img.shape[0], dst_img.shape[1]
    hi_x = max(np.max(proj[:, 0]), dw)
    hi_y = max(np.max(proj[:, 1]), dh)
    lo_x = min(np.min(proj[:, 0]), 0)
    lo_y = min(np.min(proj[:, 1]), 0)

    off_x = -lo_x if lo_x < 0 else 0
    off_y = -lo_y if lo_y < 0 else 0
    return hi_x - lo_x, hi_y - lo_y, off_x, off_y


def uniform_blend(img1, img2):
    """Overlap blend of two ``[H, W, 3]`` uint8 images on the GPU.

    Where both pixels are non-black (any channel > 0) the result is the truncated
    average, elsewhere the sum -- the integer-exact form of the reference's float64
    ``(a + b) * mask`` followed by ``astype(uint8)``.  Reference: pyviz/apap_utils.py:75-88.
    """
    from . import _runtime  # late import: the helpers above must work without a GPU

    return _runtime.blend_host(img1, img2)
