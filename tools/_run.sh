N=${1:-4}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "rc=$?"; tail -3 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${N}gpu.json').read().strip().splitlines()[-1])
c=d["c3_sharded"]; print(d['value'], {k:c[k] for k in ("dlt_ms","warp_ms","allgather_ms","warp_and_assemble_ms")})
PY
