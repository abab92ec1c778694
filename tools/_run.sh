timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01i.json 2> gpurun_out/bench_r01i.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r01i.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r01i.json').read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "roof", d["roofline"]["frac"], d["roofline"]["ms"], d["roofline"]["eig_ms"])
print("e2e dlt", d["e2e"]); print("warp", d["warp"]["ms_per_step"], d["warp"]["roofline"]["frac"], "e2e", d["warp"]["e2e"])
print("cpu", d["cpu_baseline"]); print("clocks", d["clocks"])
PY
