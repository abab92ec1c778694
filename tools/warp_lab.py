#!/usr/bin/env python
"""LAB: k_warp timing, cold (L2 flushed) vs warm (source and canvas L2-resident, no launch gap)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
p = bench.Pass(torch, dev, name)
p.gram(); p.eig(); p.prepare_warp()
flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
def run(force, warm, reps=20):
    ts = []
    for _ in range(reps + 3):
        flush.add_(1)
        if warm:
            p.st.warp_device(p.img, p.tables, p.sc.mesh_cells, out=p.canvas, force_exact=force)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        p.st.warp_device(p.img, p.tables, p.sc.mesh_cells, out=p.canvas, force_exact=force)
        e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = ts[3:]
    return sum(ts) / len(ts), min(ts)
for label, force in (("fast path", 0), ("all float64", 1)):
    for warm in (False, True):
        avg, best = run(force, warm)
        print(f"{label:16s} {'warm L2' if warm else 'cold   '} avg {avg:7.2f} us  best {best:7.2f} us")
