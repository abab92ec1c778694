timeout 600 python -m pytest tests -m gpu -x -q -k "calculate_m" -s 2>&1 | tail -8
python - <<'PY'
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from cvx_proj_b200 import spectral_method as psm
from oracle.gen_golden_spectral import OPTS, spectral_case
for name in ("s300", "s1000", "s2500"):
    c, o, cf, of, fmat, hg = spectral_case(name)
    diag = psm.affinity_diagonal(c, o, cf, of, fmat, OPTS["epi_weight"])
    psm.spectral_segment_device(c, o, diag, OPTS["affinity_eps"])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    seg, info = psm.spectral_segment_device(c, o, diag, OPTS["affinity_eps"], return_info=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{name}: N={c.shape[0]} matrix + {info['iterations']} power steps in {dt*1e3:.2f} ms")
PY
