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
from cvx_proj_b200 import spectral_method as psm, synth, utils as putils  # noqa: E402
import numpy as np  # noqa: E402

sp_o, sp_c, _h = synth.make_keypoints(1024, 768, 2500, seed=11)
sp_diag = np.full(2500, 0.95)
_rng = np.random.default_rng(0)
mq = torch.from_numpy(np.rint(_rng.uniform(0, 255, (4000, 128))).astype(np.float32)).to(dev)
mt = torch.from_numpy(np.rint(_rng.uniform(0, 255, (4000, 128))).astype(np.float32)).to(dev)

g_cw, g_ch, g_tx, g_ty, g_m = putils.warping_canvas(p.host_img.shape, p.host_img.shape, p.sc.h_gt)
g_out = torch.empty((g_ch, g_cw, 3), dtype=torch.uint8, device=dev)


def gw(mode):
    putils.warp_perspective(p.img, g_m, (g_cw, g_ch), base=p.centre, offset=(g_tx, g_ty), mode=mode, out=g_out)


for _ in range(iters):
    p.gram(); p.eig()                                   # K1, K2 one by one
    p.dlt()                                             # K1 + K2 as the public call launches them (K2 overlapped)
    p.st.condition_device(p.points[1:3].contiguous())   # k_condition, k_condition_mats (any two point sets)
    p.st.weight_bound_device(p.points[2], None, p.anchors[None])   # k_weight_bound
    p.st.kp_rows_device(p.points)                       # k_kp_rows
    p.st.kp_table_device(p.rows)                        # k_kp_blocks
    p.st.invert_grid(p.h_out.cpu().numpy().reshape(-1, 3, 3))   # k_inv_grid (+ copies)
    p.prepare_warp()                                    # k_warp_prep (+ uploads)
    p.warp(False); p.warp(True); p.blend()
    gw(0); gw(1); gw(2)                                 # k_warp_global: warp only, paste, mean blend
    if _ == 0:
        from cvx_proj_b200.utils import match_descriptors
        match_descriptors(mq, mt)                       # k_match_nn, k_match_finish
        p.st.local_homography(p.sc.src, p.sc.dst, p.sc.vertices)   # the one-call chain (k_scale_anchors first)
        psm.spectral_segment_device(sp_c, sp_o, sp_diag, 30.0, device=dev)   # k_affinity, k_power_step, k_power_diff
torch.cuda.synchronize()
print("profile pass ok", name, iters)
