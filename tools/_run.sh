timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/sweep_c4_c5.py 2>&1 | tee gpurun_out/sweep_c4_c5.jsonl
timeout 120 python tools/time_kernels.py c3 5 gram,eig
