timeout 900 python -m pytest tests -m gpu -x -q -k "c3_full or c5_keypoint or c4_batch" -s 2>&1 | tail -12
timeout 900 python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -12
