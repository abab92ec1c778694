timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_r01j.json 2> gpurun_out/bench_r01j.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r01j.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r01j.json').read().strip().splitlines()[-1])
print("value", d["value"], "stage ms", d["ms_per_step"], "k1", d["roofline"]["ms"], "k2", d["roofline"]["eig_ms"], "frac", d["roofline"]["frac"])
print("e2e dlt ms", d["e2e"]["ms_per_step"])
PY
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --c3 always > gpurun_out/bench_c3_1gpu.json 2>/dev/null; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c3_1gpu.json').read().strip().splitlines()[-1])
print(json.dumps(d["c3_sharded"]))
PY
