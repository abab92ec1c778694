#!/usr/bin/env python
"""Multi-GPU parity check over NCCL (run under torchrun, one rank per GPU):
the cell-row / row-band sharded pass, all-gathered, equals the single-GPU pass bit for bit -- H grid and panorama.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_sharded_nccl.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from cvx_proj_b200 import sharding, synth  # noqa: E402
from cvx_proj_b200.apap import APAP  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
for name, kw in (("mini", {}), ("c1", {}), ("c2", {})):
    sc = synth.make_scene(name, **kw)
    img = sc.image(1)
    st = APAP(sc.gamma, sc.sigma, [sc.final_w, sc.final_h], [sc.offset_x, sc.offset_y], device=dev)
    sh = sharding.ShardedAPAP(st, sc.mesh, sc.mesh_cells, sc.mesh_cells, rank, world)
    h_rows = sh.local_homography(sc.src, sc.dst, sc.vertices)
    band = sh.local_warp_band(img, h_rows.copy())
    pano = sh.panorama(band).cpu().numpy()
    # the same panorama assembled by the warp kernel itself: multimem.st through the NVLS multicast mapping
    sym = sharding.SymmetricPanorama(sc.final_h, sc.final_w, dev)
    fused = None
    if sym.supported:
        sym.local.fill_(7)
        dist.barrier()
        fused = sh.local_warp_panorama(img, h_rows.copy(), sym)
        torch.cuda.synchronize()
        fused = fused.cpu().numpy()
        direct = None
        if sym.multicast_ptr:                   # the warp kernel's own multimem stores (whole tile rows) as well
            sym.local.fill_(9)
            dist.barrier()
            direct = sh.local_warp_panorama(img, h_rows.copy(), sym, fused_stores=True)
            torch.cuda.synchronize()
            direct = direct.cpu().numpy()
    same_f = True if fused is None else bool(np.array_equal(fused, pano) and (direct is None or np.array_equal(direct, pano)))
    flags = torch.tensor([int(same_f)], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    same_f_all = bool(flags.item())
    # gather the H rows too (host side, via the same process group on device tensors)
    h_dev = torch.from_numpy(np.ascontiguousarray(h_rows)).to(dev)
    sizes = [s.n_cell_rows for s in sh.shards]
    pad = torch.zeros((max(sizes),) + tuple(h_dev.shape[1:]), dtype=h_dev.dtype, device=dev)
    pad[: h_dev.shape[0]] = h_dev
    allh = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(allh, pad)
    h_all = torch.cat([allh[g][: sizes[g]] for g in range(world)], 0).cpu().numpy()
    if rank == 0:
        h_one, _ = st.local_homography(sc.src, sc.dst, sc.vertices)
        one = st.local_warp(img, h_one.copy(), sc.mesh)
        same_h = np.array_equal(h_all.view(np.uint32), h_one.view(np.uint32))
        same_p = np.array_equal(pano, one)
        ok = ok and same_h and same_p and same_f_all
        print(f"{name}: {world} ranks over NCCL, H grid bit-identical to 1 GPU: {same_h}; panorama "
              f"({pano.shape[1]}x{pano.shape[0]}) bit-identical: {same_p}; panorama assembled by multicast stores "
              f"{'(no NVLS / canvas_w % 4 != 0: skipped)' if fused is None else 'identical on every rank: ' + str(same_f_all)}; bands "
              f"{[(s.px_row0, s.px_row1) for s in sh.shards]}", flush=True)
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("SHARDED NCCL CHECK", "PASSED" if ok else "FAILED", flush=True)
    sys.exit(0 if ok else 1)
