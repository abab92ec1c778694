#!/usr/bin/env python
"""The command ncu profiles for K3 alone: a few device-resident launches of the mesh warp (tile engine, then the
strip kernel), plain and fused.  usage: python tools/warp_prof.py [workload] [iterations]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
p = bench.Pass(torch, dev, name)
p.gram(); p.eig(); p.prepare_warp()
for _ in range(iters):
    for legacy in (False, True):
        for fused in (False, True):
            p.st.warp_device(p.img, p.tables, p.sc.mesh_cells, out=p.canvas, legacy=legacy,
                             centre_dev=p.centre if fused else None)
torch.cuda.synchronize()
print("warp prof ok", name, iters)
