#!/usr/bin/env python
"""LAB: end-to-end APAP.local_homography at c2 with pinned inputs (as bench.py's e2e leg), and the device time of the
one-call chain alone (CUDA events around apap_local_homography_points, inputs resident)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cvx_proj_b200 import _runtime as rt, synth
from cvx_proj_b200.apap import APAP
sc = synth.make_scene(sys.argv[1] if len(sys.argv) > 1 else "c2")
st = APAP(sc.gamma, sc.sigma, [sc.final_w, sc.final_h], [sc.offset_x, sc.offset_y])
src = rt.pinned_empty(sc.src.shape, np.float32); src[...] = sc.src
dst = rt.pinned_empty(sc.dst.shape, np.float32); dst[...] = sc.dst
ver = rt.pinned_empty(sc.vertices.shape, np.float64); ver[...] = sc.vertices
for _ in range(10):
    st.local_homography(src, dst, ver)
ts = []
for _ in range(200):
    torch.cuda.synchronize(); t0 = time.perf_counter(); st.local_homography(src, dst, ver); ts.append(time.perf_counter() - t0)
ts = np.array(ts) * 1e3
print(f"local_homography e2e: mean {ts.mean():.3f} ms  median {np.median(ts):.3f}  min {ts.min():.3f}")
# device time of the chain: time the public call's GPU work with events (copies included)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tt = []
for _ in range(50):
    torch.cuda.synchronize(); e0.record(); st.local_homography(src, dst, ver); e1.record(); e1.synchronize(); tt.append(e0.elapsed_time(e1))
print(f"GPU span of the call (first copy to last copy): median {np.median(tt):.3f} ms")
