#!/usr/bin/env python
"""Stress of the overlapped K1 -> K2 launch (programmatic dependent launch + per-tile counters): thousands of
back-to-back public-path launches at several sizes, results compared with the plain sequence every time, counters
checked to be zero at the end.  A hang here would be a deadlock of K2's spin-wait (run it under `timeout`)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from cvx_proj_b200 import synth  # noqa: E402
from cvx_proj_b200.apap import APAP, scale_anchors, weight_scale  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
for name, kw, batch in (("c1", {}, 1), ("mini", {"n_kp": 300, "mesh": 11}, 5), ("c2", {}, 1), ("c4", {}, 8)):
    sc = synth.make_scene(name, **kw)
    st = APAP(sc.gamma, sc.sigma, [sc.final_w, sc.final_h], [sc.offset_x, sc.offset_y], device=dev)
    table, tmats = st._prepare(sc.src, sc.dst)
    t = st.kp_table_device(torch.from_numpy(np.stack([table] * batch)).to(dev))
    a = torch.from_numpy(np.stack([scale_anchors(sc.vertices, weight_scale(sc.sigma))] * batch)).to(dev)
    m = torch.from_numpy(np.stack([tmats] * batch)).to(dev)
    ref = st.local_homography_device(t, a, m, batch, sc.n_cells, overlap=False).clone()
    raw = torch.from_numpy(np.stack([np.ascontiguousarray(sc.src, dtype=np.float32)] * batch)).to(dev)
    bound = st.weight_bound_device(raw, None, a)              # the clamp-free loop of K1 on odd launches
    out = torch.empty_like(ref)
    bad = 0
    t0 = time.time()
    n = iters if name != "c4" else iters // 4
    for k in range(n):
        st.local_homography_device(t, a, m, batch, sc.n_cells, out_h=out, overlap=True, t_bound=bound if k & 1 else None)
        if k % 50 == 49 or k == n - 1:
            bad += int((out.view(torch.int32) != ref.view(torch.int32)).sum().item())
    torch.cuda.synchronize()
    cnt = int(st._tile_counters(torch, dev, batch, sc.n_cells).abs().sum().item())
    print(f"{name} x{batch}: {n} overlapped launches in {time.time() - t0:.2f} s, mismatching words {bad}, counters left {cnt}",
          flush=True)
    assert bad == 0 and cnt == 0
print("STRESS PASSED")
