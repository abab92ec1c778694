#!/usr/bin/env python
"""LAB: per-tile time stamps of the tile engine's ring (APAP_TILE_LAB=4|...): for the first 64 CTAs and 32 tiles each,
0 = producer saw the stage empty, 1 = producer issued the copies, 2 = worker 0 saw the stage full, 3 = worker 0 done."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
p = bench.Pass(torch, dev, name)
p.gram(); p.eig(); p.prepare_warp()
flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    flush.add_(1)
    p.st.warp_device(p.img, p.tables, p.sc.mesh_cells, out=p.canvas)
torch.cuda.synchronize()
sc = p.st._scratch
tr = sc[sc.numel() - 64 * 32 * 4 * 8:].cpu().numpy().view(np.uint64).reshape(64, 32, 4).astype(np.int64)
t0 = tr[:, 0, 1].min()
ok = tr[:, :, 2] > 0
print(name, "lab", os.environ.get("APAP_TILE_LAB"), "tiles traced per CTA:", ok.sum(1)[:8])
lat = (tr[:, :, 2] - tr[:, :, 1])[ok]             # copies issued -> worker sees full
work = (tr[:, :, 3] - tr[:, :, 2])[ok]            # worker 0: full -> done
print(f"issue->full  median {np.median(lat):8.0f} ns  p10 {np.percentile(lat,10):8.0f}  p90 {np.percentile(lat,90):8.0f}")
print(f"full->done   median {np.median(work):8.0f} ns  p10 {np.percentile(work,10):8.0f}  p90 {np.percentile(work,90):8.0f}")
gap = (tr[:, 2:, 0] - tr[:, :-2, 3])[ok[:, 2:]]   # worker 0 done with tile k -> producer sees stage empty for k+2
print(f"done(k)->empty seen(k+2) median {np.median(gap):8.0f} ns  p90 {np.percentile(gap,90):8.0f}")
per = np.diff(tr[:, :, 2], axis=1)[ok[:, 1:]]
print(f"full(k)->full(k+1) median {np.median(per):8.0f} ns  mean {per.mean():8.0f}")
c = 5
print("CTA", c, "stamps (us since first issue):")
for k in range(12):
    print(k, [f"{(v - t0) / 1e3:7.2f}" for v in tr[c, k]])
