#!/usr/bin/env python
"""One short device-resident APAP pass (K1, K2, K3, fused K3+K4, K4), repeated a few times: the
command ncu profiles.  usage: python tools/profile_pass.py [workload] [iterations]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
p = bench.Pass(torch, dev, name)
p.gram(); p.eig()
p.prepare_warp()
for _ in range(iters):
    p.gram(); p.eig()                                   # K1, K2 one by one
    p.dlt()                                             # K1 + K2 as the public call launches them (K2 overlapped)
    p.st.kp_table_device(p.rows)                        # k_kp_blocks
    p.prepare_warp()                                    # k_warp_prep (+ uploads)
    p.warp(False); p.warp(True); p.blend()
torch.cuda.synchronize()
print("profile pass ok", name, iters)
