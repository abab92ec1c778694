#!/usr/bin/env python
"""LAB: host-side cost of local_warp's preparation on the GPU box, thread pool on/off."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cvx_proj_b200 import synth, apap as papap
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
sc = synth.make_scene("c2")
rng = np.random.default_rng(0)
h = (np.tile(sc.h_gt.astype(np.float32), (200, 200, 1, 1)) * (1 + 1e-3 * rng.standard_normal((200, 200, 3, 3)))).astype(np.float32)
col, row = papap.cell_lookup_tables(sc.mesh, sc.final_w, sc.final_h, 200, 200)
orig = papap._map_row_chunks
for mode in ("pool", "serial"):
    if mode == "serial":
        papap._map_row_chunks = lambda fn, n, min_rows=16: [fn(0, n)]
    best = [1e9, 1e9]
    for _ in range(5):
        g = h.copy()
        t0 = time.perf_counter(); papap.invert_grid_inplace(g); t1 = time.perf_counter()
        papap.build_warp_tables(g, col, row, sc.offset_x, sc.offset_y, sc.width, sc.height); t2 = time.perf_counter()
        best = [min(best[0], t1 - t0), min(best[1], t2 - t1)]
    print(f"{mode:7s} inv {best[0]*1e3:6.2f} ms  tables {best[1]*1e3:6.2f} ms")
papap._map_row_chunks = orig
