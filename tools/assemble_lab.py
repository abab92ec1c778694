#!/usr/bin/env python
"""LAB (torchrun, one rank per GPU): where the time of the sharded c3 "warp + assemble the panorama on every GPU" goes.
Every variant starts from a group barrier (ranks aligned, L2 flushed) and ends after the group barrier that makes the
panorama complete everywhere; device-timed, max over ranks of the mean.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/assemble_lab.py
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from cvx_proj_b200 import _runtime as rt, sharding, synth  # noqa: E402
from cvx_proj_b200.apap import cell_lookup_tables  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
cfg = synth.CONFIGS[name]
sc0 = synth.make_scene(name)
col, row = cell_lookup_tables(sc0.mesh, sc0.final_w, sc0.final_h, cfg["mesh"], cfg["mesh"])
shards = sharding.plan_shards(row, cfg["mesh"], world)
me = shards[rank]
p = bench.Pass(torch, dev, name, seed=0, rows=(me.cell_row0, me.cell_row1))
p.gram(); p.eig()
p.prepare_warp(px_rows=(me.px_row0, me.px_row1))
sym = sharding.SymmetricPanorama(sc0.final_h, sc0.final_w, dev)
assert sym.supported, "no NVLS multicast"
own = sym.local[me.px_row0:me.px_row1]
lib = rt.load_library()
stream = rt.stream_ptr(torch, dev)
off = me.px_row0 * sc0.final_w * 3
order = [(rank + 1 + k) % world for k in range(world - 1)]          # every rank starts with a different peer
peer_ptrs = (ctypes.c_void_p * len(order))(*[int(sym.handle.buffer_ptrs[r]) + off for r in order])
peers = [sym.handle.get_buffer(r, tuple(sym.local.shape), torch.uint8) for r in range(world)]
flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)


def warp_local():
    p.st.warp_device(p.img, p.tables, p.sc.mesh_cells, out=own)


def v_barrier():
    pass


def v_warp_only():
    warp_local()


def v_direct():
    p.st.warp_device(p.img, p.tables, p.sc.mesh_cells, multicast_ptr=sym.band_ptr(me.px_row0))


def v_mc_copy():
    warp_local()
    rt.check(lib.apap_multicast_copy(own.data_ptr(), sym.band_ptr(me.px_row0), own.numel(), stream), "mc")


def v_mc_copy_alone():
    rt.check(lib.apap_multicast_copy(own.data_ptr(), sym.band_ptr(me.px_row0), own.numel(), stream), "mc")


def v_peer_copy():
    warp_local()
    rt.check(lib.apap_peer_copy(own.data_ptr(), peer_ptrs, len(order), own.numel(), stream), "peer")


def v_peer_copy_alone():
    rt.check(lib.apap_peer_copy(own.data_ptr(), peer_ptrs, len(order), own.numel(), stream), "peer")


def v_copy_engines():
    warp_local()
    for r in order:
        peers[r][me.px_row0:me.px_row1].copy_(own, non_blocking=True)


def timed(body, reps=20, warm=4):
    ts = []
    for k in range(reps + warm):
        flush.add_(1)
        sym.barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        body()
        sym.barrier()
        e1.record()
        e1.synchronize()
        if k >= warm:
            ts.append(e0.elapsed_time(e1) * 1e3)
    t = torch.tensor([sum(ts) / len(ts), min(ts)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def checksum(t):
    v = t.reshape(-1).to(torch.int64)
    return (v * (torch.arange(v.numel(), device=dev, dtype=torch.int64) % 8191 + 1)).sum()


ref = torch.empty_like(own)
p.st.warp_device(p.img, p.tables, p.sc.mesh_cells, out=ref)
sums = [torch.zeros((), dtype=torch.int64, device=dev) for _ in range(world)]
dist.all_gather(sums, checksum(ref))
if rank == 0:
    print(f"{name}: {world} GPUs, panorama {sc0.final_w}x{sc0.final_h} = {3 * sc0.canvas_px / 1e6:.1f} MB, band "
          f"{own.numel() / 1e6:.1f} MB", flush=True)
for label, body, check in (("barrier only", v_barrier, False), ("warp only (+barrier)", v_warp_only, False),
                           ("warp kernel's own multimem stores", v_direct, True),
                           ("warp + multicast copy kernel", v_mc_copy, True),
                           ("multicast copy kernel alone", v_mc_copy_alone, False),
                           ("warp + unicast peer copy kernel", v_peer_copy, True),
                           ("unicast peer copy kernel alone", v_peer_copy_alone, False),
                           ("warp + copy engines (N-1 peer memcpy)", v_copy_engines, True)):
    if check:
        sym.local.fill_(3)
        torch.cuda.synchronize(); dist.barrier()
    avg, best = timed(body)
    okay = ""
    if check:
        torch.cuda.synchronize(); dist.barrier()
        rows_ok = [bool(checksum(sym.local[s.px_row0:s.px_row1]) == sums[s.rank]) for s in shards]
        flag = torch.tensor([int(all(rows_ok))], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        okay = f"  panorama complete on every rank: {bool(flag.item())}"
    if rank == 0:
        print(f"{label:42s} avg {avg:8.2f} us  best {best:8.2f} us{okay}", flush=True)
dist.barrier()
dist.destroy_process_group()
