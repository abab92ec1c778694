#!/usr/bin/env python
"""LAB: one small mesh warp (scene name from argv) checked against the oracle -- the command compute-sanitizer runs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cvx_proj_b200 import synth
from cvx_proj_b200.apap import APAP
from oracle import apap_oracle as orc
name = sys.argv[1] if len(sys.argv) > 1 else "tiny"
sc = synth.make_scene(name)
img = sc.image(1)
st = APAP(sc.gamma, sc.sigma, [sc.final_w, sc.final_h], [sc.offset_x, sc.offset_y])
h = orc.local_homography_gram64(sc.src, sc.dst, sc.vertices, sc.gamma, sc.sigma)
print(name, "src", img.shape, "canvas", sc.final_w, sc.final_h, flush=True)
got = st.local_warp(img, h.copy(), sc.mesh)
want = orc.local_warp(img, orc.invert_grid(h), sc.mesh, (sc.final_w, sc.final_h), (sc.offset_x, sc.offset_y))
print("equal:", np.array_equal(got, want))
