timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 120 python tools/time_kernels.py c2 10 gram,eig
timeout 120 python tools/time_kernels.py c3 5 gram,eig
