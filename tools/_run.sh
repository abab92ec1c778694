timeout 600 python -m pytest tests -m gpu -x -q -k "image_warping or warp_perspective" 2>&1 | tail -5
timeout 120 python tools/time_kernels.py c2 10 gwarp,gwarp_paste,gwarp_mean,warp
timeout 120 python tools/time_kernels.py c3 5 gwarp,gwarp_paste,gwarp_mean
