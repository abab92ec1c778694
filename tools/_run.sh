timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/check_sharded_nccl.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tee gpurun_out/sharded_nccl_check.txt
timeout 300 python -m pytest tests -m gpu -x -q -k "warp_batch" 2>&1 | tail -3
