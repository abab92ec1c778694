"""Property tests (hypothesis) of the host-side logic that feeds the kernels: the reference's lookup rule,
the row-block work list, the tensor-core table split, the canvas arithmetic of the global warp.  CPU only."""
import numpy as np
from hypothesis import given, settings, strategies as st

from cvx_proj_b200 import apap as papap
from cvx_proj_b200 import apap_utils, utils as putils
from cvx_proj_b200 import _runtime as rt

SETTINGS = dict(max_examples=60, deadline=None)


@settings(**SETTINGS)
@given(w=st.integers(1, 400), h=st.integers(1, 300), mesh=st.integers(1, 40))
def test_cell_lookup_is_the_references_np_where_rule(w, h, mesh):
    """pyviz/apap.py:207,209: the cell of pixel k is np.where(k < edges)[0][0] - 1, for get_mesh's linspace edges."""
    edges = apap_utils.get_mesh((w, h), mesh + 1)
    col, row = papap.cell_lookup_tables(edges, w, h, mesh, mesh)
    for lut, e, extent in ((col, edges[0], w), (row, edges[1], h)):
        want = np.array([np.where(k < e)[0][0] - 1 for k in range(extent)])
        assert np.array_equal(lut.astype(np.int64), np.mod(want, mesh))
        assert (np.diff(lut.astype(np.int64)) >= 0).all()          # monotone for a mesh that starts at 0


@settings(**SETTINGS)
@given(lens=st.lists(st.integers(1, 23), min_size=1, max_size=12), data=st.data())
def test_row_blocks_cover_every_band_row_and_never_cross_a_cell_row(lens, data):
    row = np.repeat(np.arange(len(lens)), lens).astype(np.uint16)
    first = np.r_[0, np.cumsum(lens)[:-1]]
    fh = row.shape[0]
    r0 = data.draw(st.integers(0, fh))
    r1 = data.draw(st.integers(r0, fh))
    blocks = papap.build_row_blocks(row, first, r0, r1)
    seen = np.zeros(fh, dtype=np.int64)
    prev = -1
    for w0, w1 in blocks:
        i0, n, cr, dy0 = int(w0) & 0x0fffffff, int(w0) >> 28, int(w1) & 0xffff, int(w1) >> 16
        assert 1 <= n <= rt.WARP_BLOCK_ROWS and i0 > prev
        prev = i0
        assert r0 <= i0 and i0 + n <= r1
        assert (row[i0:i0 + n] == cr).all() and dy0 == i0 - first[cr]
        seen[i0:i0 + n] += 1
    assert (seen[r0:r1] >= 1).all() and not seen[:r0].any() and not seen[r1:].any()
    assert seen.max(initial=0) <= 2                                  # a row is computed at most twice (overlap of full blocks)


@settings(**SETTINGS)
@given(w=st.integers(2, 300), mesh=st.integers(1, 60))
def test_warp_luts_extents_match_the_lookup(w, mesh):
    edges = apap_utils.get_mesh((w, w), mesh + 1)
    col, row = papap.cell_lookup_tables(edges, w, w, mesh, mesh)
    col_lut, row_first, col_ext, row_ext = papap.warp_luts(col, row, mesh, mesh)
    for c in range(mesh):
        px = np.flatnonzero(col == c)
        if px.size:
            assert tuple(col_ext[c]) == (px[0], px[-1])
            dx = col_lut[px, 1].view(np.float32)
            assert np.array_equal(dx, (px - px[0]).astype(np.float32)) and (col_lut[px, 0] == c).all()
        else:
            assert col_ext[c, 0] > col_ext[c, 1]
    assert np.array_equal(row_first[row], np.array([np.flatnonzero(row == r)[0] for r in row]))


@settings(**SETTINGS)
@given(st.lists(st.floats(-1e4, 1e4, allow_nan=False, width=32), min_size=24, max_size=24), st.integers(1, 20))
def test_kp_blocks_split_is_exact_and_tf32(vals, n_rows):
    """P = Ph + Pl exactly, Ph has at most 11 significant bits (TF32), every element lands where the MMA's K-major
    core-matrix layout expects it."""
    table = np.zeros((128, rt.KP_ROW), dtype=np.float32)
    table[:n_rows, :24] = np.array(vals, dtype=np.float32)[None, :] * (1 + np.arange(n_rows, dtype=np.float32)[:, None])
    table[:, 24] = table[:, 25] = np.arange(128)
    table[:, 26] = table[:, 27] = -np.arange(128)
    blk = papap.build_kp_blocks(table)
    assert blk.shape == (16, rt.KP_BLOCK_FLOATS)
    for kb in range(16):
        for k in range(8):
            for n in (0, 7, 23):
                off = (k // 4) * 256 + (n // 8) * 32 + (n % 8) * 4 + k % 4
                hi, lo = blk[kb, off], blk[kb, (k // 4) * 256 + ((32 + n) // 8) * 32 + ((32 + n) % 8) * 4 + k % 4]
                p = table[kb * 8 + k, n]
                assert np.float32(hi) + np.float32(lo) == p
                assert (np.float32(hi).view(np.uint32) & np.uint32(0x1FFF)) == 0
        assert np.array_equal(blk[kb, 512:520], table[kb * 8:kb * 8 + 8, 24])
        assert np.array_equal(blk[kb, 520:528], table[kb * 8:kb * 8 + 8, 26])


@settings(**SETTINGS)
@given(w=st.integers(8, 500), h=st.integers(8, 400), dx=st.floats(-200, 200), dy=st.floats(-200, 200),
       s=st.floats(0.5, 2.0), p=st.floats(-1e-4, 1e-4))
def test_warping_canvas_holds_both_images(w, h, dx, dy, s, p):
    """pyviz/utils.py:99-112: the canvas contains the base image at the offsets and the projected corners."""
    hmat = np.array([[s, 0.05, dx], [-0.03, s, dy], [p, -p, 1.0]])
    cw, ch, tx, ty, m = putils.warping_canvas((h, w, 3), (h, w, 3), hmat)
    assert tx >= 0 and ty >= 0 and tx + w <= cw and ty + h <= ch
    corners = np.array([[0, 0, 1], [0, h, 1], [w, h, 1], [w, 0, 1]], dtype=np.float64) @ m.T
    xy = corners[:, :2] / corners[:, 2:3]
    assert (xy > -1.0).all() and (xy[:, 0] < cw + 1.0).all() and (xy[:, 1] < ch + 1.0).all()
    inv = putils.invert3x3(m)
    assert np.allclose(inv @ m, np.eye(3), atol=1e-9)
