"""Multi-GPU partition of one APAP pass: cell rows and the canvas row bands they own.

Cells are independent (pyviz/apap.py:147-168 carries nothing between iterations) and the
pixel -> cell lookup is by canvas row (pyviz/apap.py:207), so rank g that owns the cell rows
``[m0, m1)`` can compute those cells' homographies AND warp exactly the canvas rows whose
``row_cell`` lies in ``[m0, m1)`` with no exchange in between (keypoints and the source image
are replicated).  The only collective is the all-gather that assembles the panorama from the
row bands (NCCL over NVLink on GPUs; the same code runs over gloo on CPU tensors in tests).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Shard:
    rank: int
    cell_row0: int      # first owned cell row
    cell_row1: int      # one past the last owned cell row
    px_row0: int        # first canvas row of the band
    px_row1: int        # one past the last canvas row of the band

    @property
    def n_cell_rows(self) -> int:
        return self.cell_row1 - self.cell_row0

    @property
    def n_px_rows(self) -> int:
        return self.px_row1 - self.px_row0


def split_rows(n_rows: int, world: int):
    """Contiguous, balanced ``[(r0, r1)] * world`` (the first ``n_rows % world`` ranks get one more)."""
    base, extra = divmod(n_rows, world)
    out, r = [], 0
    for g in range(world):
        n = base + (1 if g < extra else 0)
        out.append((r, r + n))
        r += n
    return out


def plan_shards(row_cell: np.ndarray, grid_rows: int, world: int):
    """One ``Shard`` per rank from the canvas-row -> cell-row table of the warp.

    ``row_cell`` must be non-decreasing (true for any edge array from ``get_mesh``); the bands
    then tile the canvas rows exactly once.
    """
    row_cell = np.asarray(row_cell, dtype=np.int64)
    if row_cell.size and np.any(np.diff(row_cell) < 0):
        raise ValueError("row_cell must be non-decreasing to shard the canvas by cell row")
    shards = []
    for g, (m0, m1) in enumerate(split_rows(grid_rows, world)):
        r0 = int(np.searchsorted(row_cell, m0, side="left"))
        r1 = int(np.searchsorted(row_cell, m1, side="left"))
        shards.append(Shard(g, m0, m1, r0, r1))
    return shards


def gather_bands(band, shards, canvas_w: int, group=None):
    """All-gather the row bands into the full ``[canvas_h, canvas_w, 3]`` uint8 panorama.

    ``band`` is this rank's ``[n_px_rows, canvas_w, 3]`` uint8 tensor (CUDA -> NCCL, CPU -> gloo).
    Bands differ by at most a few rows, so each is padded to the tallest band for one
    ``all_gather_into_tensor`` and the padding is dropped when the rows are stitched.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    assert world == len(shards)
    tallest = max(s.n_px_rows for s in shards)
    row_bytes = canvas_w * 3
    mine = torch.zeros((tallest, row_bytes), dtype=torch.uint8, device=band.device)
    mine[: band.shape[0]] = band.reshape(band.shape[0], row_bytes)
    everyone = torch.empty((world, tallest, row_bytes), dtype=torch.uint8, device=band.device)
    dist.all_gather_into_tensor(everyone.view(-1), mine.view(-1), group=group)
    parts = [everyone[s.rank, : s.n_px_rows] for s in shards]
    return torch.cat(parts, dim=0).reshape(-1, canvas_w, 3)


class SymmetricPanorama:
    """The panorama buffer of a sharded pass as torch symmetric memory: one ``[canvas_h, canvas_w, 3]`` uint8 buffer
    per GPU, mapped into every peer, plus the NVLS multicast mapping of all of them.  The warp kernel stores its row
    band through the multicast address (``multimem.st``): the NVSwitch delivers it to every GPU, so when the group has
    passed ``barrier()`` each rank holds the complete panorama -- the assembly is fused into the warp and there is
    no all-gather.  Needs NVSwitch multicast (``multicast_ptr`` non-zero) and ``canvas_w % 16 == 0`` (16-byte stores)."""

    def __init__(self, canvas_h: int, canvas_w: int, device, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group if group is not None else dist.group.WORLD
        self.canvas_h, self.canvas_w = int(canvas_h), int(canvas_w)
        self.local = symm_mem.empty((self.canvas_h, self.canvas_w, 3), dtype=torch.uint8, device=device)
        self.handle = symm_mem.rendezvous(self.local, self.group)
        self.multicast_ptr = int(self.handle.multicast_ptr or 0)
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self._peers = None

    @property
    def supported(self) -> bool:
        return self.multicast_ptr != 0 and self.canvas_w % 16 == 0

    def band_ptr(self, px_row0: int) -> int:
        return self.multicast_ptr + int(px_row0) * self.canvas_w * 3

    def broadcast_band(self, px_row0: int, px_row1: int):
        """Store the canvas rows ``[px_row0, px_row1)`` of this rank's panorama into every GPU's (``apap_multicast_copy``)."""
        import torch

        from . import _runtime as rt

        band = self.local[px_row0:px_row1]
        if band.numel() == 0:
            return
        if self.world != 2 and not self.supported:
            raise rt.ApapError("broadcast_band: this group has no NVLS multicast mapping (multicast_ptr == 0) or "
                               "canvas_w % 16 != 0 -- assemble the panorama with gather_bands instead")
        if self.world == 2:
            # two GPUs: one unicast peer copy beats the multicast store rate (measured 129 us against 190 us for
            # the 62 MB band of c3); from three GPUs on a rank would send its band N - 1 times and multicast wins
            if self._peers is None:
                self._peers = [self.handle.get_buffer(r, tuple(self.local.shape), torch.uint8) for r in range(self.world)]
            self._peers[1 - self.rank][px_row0:px_row1].copy_(band, non_blocking=True)
            return
        with torch.cuda.device(self.local.device):
            rt.check(rt.load_library().apap_multicast_copy(band.data_ptr(), self.band_ptr(px_row0), band.numel(),
                                                           rt.stream_ptr(torch, self.local.device)), "apap_multicast_copy")

    def barrier(self):
        """Every rank's band has landed everywhere once all ranks are past this (device-side, on the current stream)."""
        self.handle.barrier()


def _image_on_device(rt, torch, device, img):
    """uint8, contiguous, on ``device`` -- what ``APAP._warp`` does with its image argument."""
    if isinstance(img, np.ndarray):
        return rt.to_device(torch, device, img.astype(np.uint8, copy=False))
    if img.dtype != torch.uint8:
        raise TypeError("a device image must be a uint8 tensor")
    return img.contiguous()


class ShardedAPAP:
    """One APAP pass split over the ranks of a ``torch.distributed`` group.

    Every rank holds the full keypoint set and source image; it solves its cell rows (K1 + K2),
    inverts them on the host like the reference, and warps its row band (K3).  ``panorama()``
    is the one collective.
    """

    def __init__(self, stitcher, mesh, grid_rows: int, grid_cols: int, rank: int, world: int):
        from .apap import cell_lookup_tables

        self.stitcher = stitcher
        self.mesh = np.asarray(mesh)
        self.grid_rows, self.grid_cols = grid_rows, grid_cols
        self.col_cell, self.row_cell = cell_lookup_tables(self.mesh, int(stitcher.final_width),
                                                          int(stitcher.final_height), grid_rows, grid_cols)
        self.shards = plan_shards(self.row_cell, grid_rows, world)
        self.me = self.shards[rank]

    def local_homography(self, src_point, dst_point, vertices):
        """H for the owned cell rows only: ``[n_cell_rows, grid_cols, 3, 3]`` float32."""
        s = self.me
        h, _ = self.stitcher.local_homography(src_point, dst_point, np.asarray(vertices)[s.cell_row0:s.cell_row1])
        return h

    def local_warp_band(self, ori_img, local_h_rows):
        """Warp the owned canvas rows.  ``local_h_rows`` = this rank's rows of H (inverted in place,
        like ``APAP.local_warp``).  Returns a device tensor ``[n_px_rows, canvas_w, 3]``."""
        from . import _runtime as rt

        st, s = self.stitcher, self.me
        torch, device = rt.torch_cuda(st.device)
        st.invert_grid(local_h_rows, device)
        full = np.zeros((self.grid_rows, self.grid_cols, 3, 3), dtype=np.float32)
        full[...] = np.eye(3, dtype=np.float32)
        full[s.cell_row0:s.cell_row1] = local_h_rows
        src_dev = _image_on_device(rt, torch, device, ori_img)
        tables = st.warp_tables_device(full, self.col_cell, self.row_cell, int(ori_img.shape[1]),
                                       int(ori_img.shape[0]), device, s.px_row0, s.px_row1)
        return st.warp_device(src_dev, tables, self.grid_cols)

    def panorama(self, band, group=None):
        return gather_bands(band, self.shards, int(self.stitcher.final_width), group)

    def local_warp_panorama(self, ori_img, local_h_rows, pano: "SymmetricPanorama", fused_stores: bool = False):
        """``local_warp_band`` + ``panorama`` without an all-gather: the owned canvas rows are warped into this rank's
        panorama and broadcast into every other GPU's through the NVLS multicast mapping (``fused_stores``: by the
        warp kernel's own stores -- whole 384-byte tile rows out of shared memory -- instead of a broadcast kernel);
        a group barrier on entry (peers may still read the previous pass) and on exit; returns this rank's
        (complete) panorama tensor.  Raises ``ApapError`` when the group has no multicast mapping."""
        from . import _runtime as rt

        st, s = self.stitcher, self.me
        if (fused_stores or pano.world != 2) and not pano.supported:
            # a bare offset would pass for a multicast address and multimem.st would fault the GPU
            raise rt.ApapError("local_warp_panorama: this group has no NVLS multicast mapping (multicast_ptr == 0) or "
                               "canvas_w % 16 != 0 -- use local_warp_band + panorama (all-gather) instead")
        torch, device = rt.torch_cuda(st.device)
        # no rank may store into a peer's panorama while that peer still reads the previous pass
        pano.barrier()
        st.invert_grid(local_h_rows, device)
        full = np.zeros((self.grid_rows, self.grid_cols, 3, 3), dtype=np.float32)
        full[...] = np.eye(3, dtype=np.float32)
        full[s.cell_row0:s.cell_row1] = local_h_rows
        src_dev = _image_on_device(rt, torch, device, ori_img)
        tables = st.warp_tables_device(full, self.col_cell, self.row_cell, int(ori_img.shape[1]),
                                       int(ori_img.shape[0]), device, s.px_row0, s.px_row1)
        if fused_stores:
            st.warp_device(src_dev, tables, self.grid_cols, multicast_ptr=pano.band_ptr(s.px_row0))
        else:
            band = pano.local[s.px_row0:s.px_row1]
            st.warp_device(src_dev, tables, self.grid_cols, out=band)
            pano.broadcast_band(s.px_row0, s.px_row1)
        pano.barrier()
        return pano.local
