timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01n.json 2> gpurun_out/bench_r01n.err; echo "rc=$?"; tail -3 gpurun_out/bench_r01n.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r01n.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['spectral'], d['global_warp']['ms_per_step'], d['gpu_launches'])
PY
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
