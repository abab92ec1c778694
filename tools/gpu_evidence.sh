#!/bin/bash
# One gpurun call that collects the round's evidence: GPU tests, both bench arms, the ncu launch list
# and one `--set full` capture of every kernel of the pass.  usage: tools/gpu_evidence.sh <tag>
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"
tail -3 $out/pytest_gpu_$tag.log
python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err; echo "ref rc=$?"
python tools/profile_pass.py c2 2 > $out/pp_plain_$tag.log 2>&1; echo "profile_pass rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $out/launches_$tag.csv \
    python tools/profile_pass.py c2 2 > $out/ncu_list_$tag.log 2>&1; echo "ncu list rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/launches_bench_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu --c3 multi > $out/ncu_list_bench_$tag.log 2>&1; echo "ncu bench list rc=$?"
# one `--set full` capture of every kernel of the pass (kept under 64 MiB: gpurun does not bring back more)
if [ "${2:-full}" = full ]; then
ncu --set full --clock-control none --import-source on -k regex:'k_gram|k_eig|k_warp|k_tile_prep|k_blend|k_kp_rows|k_kp_blocks|k_condition|k_inv_grid|k_affinity|k_power_step|k_power_diff' -c 36 \
    -o $out/prof_$tag -f python tools/profile_pass.py c2 1 > $out/ncu_full_$tag.log 2>&1; echo "ncu full rc=$?"
fi
ncu --set full --clock-control none -k regex:'k_match|k_scale_anchors|k_weight_bound' -c 6 \
    -o $out/prof_${tag}_new -f python tools/profile_pass.py c2 1 > $out/ncu_full_${tag}_new.log 2>&1; echo "ncu new kernels rc=$?"
head -c 1500 $out/bench_$tag.json
