"""Where the end-to-end time of the public calls goes (host side): cProfile of ``APAP.local_homography`` and
``APAP.local_warp`` at a bench configuration, pinned numpy inputs.  Usage: python tools/profile_e2e.py [config] [calls]"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvx_proj_b200 import _runtime as rt  # noqa: E402
from cvx_proj_b200.apap import APAP  # noqa: E402
from cvx_proj_b200 import synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 100
sc = synth.make_scene(name)
st = APAP(sc.gamma, sc.sigma, [sc.final_w, sc.final_h], [sc.offset_x, sc.offset_y])
src = rt.pinned_empty(sc.src.shape, np.float32); src[...] = sc.src
dst = rt.pinned_empty(sc.dst.shape, np.float32); dst[...] = sc.dst
img0 = sc.image(1)
img = rt.pinned_empty(img0.shape, np.uint8); img[...] = img0
for _ in range(5):
    h, _ = st.local_homography(src, dst, sc.vertices)
    st.local_warp(img, h, sc.mesh)
for label, fn in (("local_homography", lambda: st.local_homography(src, dst, sc.vertices)),
                  ("local_warp", lambda: st.local_warp(img, h, sc.mesh))):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(calls):
        fn()
    wall = (time.perf_counter() - t0) / calls
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(calls):
        fn()
    pr.disable()
    print(f"==== {label} @ {name}: {wall * 1e3:.3f} ms per call (unprofiled), {calls} calls under cProfile")
    pstats.Stats(pr).sort_stats("tottime").print_stats(16)
