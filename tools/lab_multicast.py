"""Lab (torchrun, >= 2 GPUs): where the time of the multicast panorama goes -- the warp kernel with multimem stores,
the group barrier, a plain copy of the band into the multicast mapping, and NCCL's all-gather of the bands."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from cvx_proj_b200 import sharding, synth
from cvx_proj_b200.apap import cell_lookup_tables

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = synth.CONFIGS["c3"]; sc0 = synth.make_scene("c3")
col, row = cell_lookup_tables(sc0.mesh, sc0.final_w, sc0.final_h, cfg["mesh"], cfg["mesh"])
shards = sharding.plan_shards(row, cfg["mesh"], world); me = shards[rank]
p = bench.Pass(torch, dev, "c3", seed=0, rows=(me.cell_row0, me.cell_row1))
p.gram(); p.eig(); p.prepare_warp(px_rows=(me.px_row0, me.px_row1))
sym = sharding.SymmetricPanorama(sc0.final_h, sc0.final_w, dev)
flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=10):
    ts = []
    for k in range(n + 3):
        dist.barrier(); flush.add_(1)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        if k >= 3: ts.append(e0.elapsed_time(e1) * 1e3)
    t = torch.tensor([float(np.median(ts))], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


mc_band = sym.band_ptr(me.px_row0)
res = {
    "warp local": timed(lambda: p.warp(False)),
    "warp multicast (no barrier)": timed(lambda: p.st.warp_device(p.img, p.tables, p.sc.mesh_cells, multicast_ptr=mc_band)),
    "barrier only": timed(lambda: sym.barrier()),
    "warp multicast + barrier": timed(lambda: (p.st.warp_device(p.img, p.tables, p.sc.mesh_cells, multicast_ptr=mc_band), sym.barrier())),
    "nccl all-gather of the bands": timed(lambda: sharding.gather_bands(p.canvas, shards, sc0.final_w)),
    "multicast broadcast kernel of the band": timed(lambda: sym.broadcast_band(me.px_row0, me.px_row1)),
    "warp local + broadcast kernel + barrier": timed(lambda: (p.st.warp_device(p.img, p.tables, p.sc.mesh_cells, out=sym.local[me.px_row0:me.px_row1]), sym.broadcast_band(me.px_row0, me.px_row1), sym.barrier())),
}
# plain copy of the local band into every peer's buffer through P2P pointers (copy engine / SM copy)
peers = [sym.handle.get_buffer(r, (sc0.final_h, sc0.final_w, 3), torch.uint8) for r in range(world)]
def p2p_copy():
    for r in range(world):
        peers[r][me.px_row0:me.px_row1].copy_(p.canvas, non_blocking=True)
res["band copied to every peer with tensor.copy_ (P2P)"] = timed(p2p_copy)
if rank == 0:
    for k, v in res.items():
        print(f"{world} GPUs, c3 band {p.canvas.numel() / 1e6:.1f} MB: {k:55s} {v:8.1f} us", flush=True)
dist.barrier(); dist.destroy_process_group()
