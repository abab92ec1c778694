#!/usr/bin/env python
"""LAB: time the tile engine alone (cold L2), optionally under APAP_TILE_LAB timing experiments (wrong pixels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
p = bench.Pass(torch, dev, name)
p.gram(); p.eig(); p.prepare_warp()
flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for _ in range(23):
    flush.add_(1)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    p.st.warp_device(p.img, p.tables, p.sc.mesh_cells, out=p.canvas)
    e1.record(); e1.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
ts = ts[3:]
print(f"{name} lab={os.environ.get('APAP_TILE_LAB', '0')} avg {sum(ts)/len(ts):7.2f} us best {min(ts):7.2f} us")
