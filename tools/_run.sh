timeout 600 python tools/stress_overlap.py 4000 2>&1 | tee gpurun_out/stress_overlap.txt
