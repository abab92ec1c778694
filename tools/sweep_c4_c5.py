#!/usr/bin/env python
"""BASELINE configs c4 and c5 on one B200 (device-resident, CUDA events, L2 flushed before every launch):
  c5: keypoint sweep N = 1k .. 64k at a fixed 256 x 256 grid -- where K1 goes from latency- to XU-bound
  c4: a batch of 64 1080p pairs (2k keypoints, 100 x 100 grid) in ONE launch of K1 + K2 (grid.z = pair)
Prints one JSON line per measurement (copied to profiles/)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from cvx_proj_b200 import _runtime as rt, synth  # noqa: E402
from cvx_proj_b200.apap import APAP, scale_anchors, weight_scale  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
mufu_peak = rt.pipe_peak(rt.PROBE_MUFU, dev)


def timed(fn, iters=10, warm=3):
    ts = []
    for k in range(iters + warm):
        flush.add_(1)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        if k >= warm:
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def scene_tensors(sc, st):
    table, tmats = st._prepare(sc.src, sc.dst)
    return table, tmats, scale_anchors(sc.vertices, weight_scale(sc.sigma))


# ---- c5
for n_kp in (1000, 2000, 4000, 8000, 16000, 32000, 64000):
    sc = synth.make_scene("c5", n_kp=n_kp)
    st = APAP(sc.gamma, sc.sigma, [sc.final_w, sc.final_h], [sc.offset_x, sc.offset_y], device=dev)
    table, tmats, anchors = scene_tensors(sc, st)
    t_dev = st.kp_table_device(torch.from_numpy(table[None]).to(dev))
    a_dev = torch.from_numpy(anchors[None]).to(dev)
    m_dev = torch.from_numpy(tmats[None]).to(dev)
    cells, n_pad = sc.n_cells, table.shape[0]
    ks, cp, nbytes = rt.gram_plan(cells, n_pad, rt.GRAM_TCGEN05)
    partials = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
    out_h = torch.empty((1, cells, 9), dtype=torch.float32, device=dev)
    lib, s = rt.load_library(), rt.stream_ptr(torch, dev)
    g2 = float(np.float32(sc.gamma ** 2))
    ms_gram = timed(lambda: rt.check(lib.apap_gram_partials(t_dev.data_ptr(), a_dev.data_ptr(), 1, cells, n_pad, g2,
                                                            rt.GRAM_TCGEN05, None, partials.data_ptr(), s)))
    ms_eig = timed(lambda: rt.check(lib.apap_eig_denorm(partials.data_ptr(), m_dev.data_ptr(), 1, cells, ks,
                                                        rt.EIG_AUTO, out_h.data_ptr(), None, s)))
    ms_both = timed(lambda: st.local_homography_device(t_dev, a_dev, m_dev, 1, cells, out_h=out_h, partials=partials))
    # spot check against the float64 Gram oracle-free invariant: finite and H[2][2] == 1
    h = out_h.cpu().numpy()
    assert np.isfinite(h).all() and np.allclose(h[..., 8], 1.0)
    print(json.dumps({"config": "c5", "n_kp": n_kp, "n_kp_padded": n_pad, "cells": cells, "k_splits": ks,
                      "gram_ms": ms_gram, "eig_ms": ms_eig, "k1_k2_ms": ms_both, "cells_per_s": cells / (ms_both * 1e-3),
                      "xu_frac": 2.0 * cells * n_pad / (ms_gram * 1e-3) / mufu_peak,
                      "partials_mb": nbytes / 1e6}), flush=True)

# ---- c4
pairs = 64
scs = [synth.make_scene("c4", seed=k) for k in range(pairs)]
st = APAP(scs[0].gamma, scs[0].sigma, [scs[0].final_w, scs[0].final_h], [scs[0].offset_x, scs[0].offset_y], device=dev)
prep = [scene_tensors(sc, st) for sc in scs]
rows = torch.from_numpy(np.stack([p[0] for p in prep])).to(dev)
t_dev = st.kp_table_device(rows)
m_dev = torch.from_numpy(np.stack([p[1] for p in prep])).to(dev)
a_dev = torch.from_numpy(np.stack([p[2] for p in prep])).to(dev)
cells, n_pad = scs[0].n_cells, rows.shape[1]
ks, cp, nbytes = rt.gram_plan(cells, n_pad, rt.GRAM_TCGEN05)
partials = torch.empty(pairs * nbytes // 4, dtype=torch.float32, device=dev)
out_h = torch.empty((pairs, cells, 9), dtype=torch.float32, device=dev)
ms_batch = timed(lambda: st.local_homography_device(t_dev, a_dev, m_dev, pairs, cells, out_h=out_h, partials=partials))
ms_single = timed(lambda: st.local_homography_device(t_dev[:1], a_dev[:1], m_dev[:1], 1, cells, out_h=out_h[:1],
                                                     partials=partials))
print(json.dumps({"config": "c4", "pairs": pairs, "cells_per_pair": cells, "n_kp_padded": n_pad, "k_splits": ks,
                  "batch_ms": ms_batch, "cells_per_s": pairs * cells / (ms_batch * 1e-3),
                  "single_pair_ms": ms_single, "single_pair_cells_per_s": cells / (ms_single * 1e-3),
                  "xu_frac_batch": 2.0 * pairs * cells * n_pad / (ms_batch * 1e-3) / mufu_peak}), flush=True)
