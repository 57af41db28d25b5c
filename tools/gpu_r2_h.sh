#!/bin/bash
# 2 GPUs: the C ABI multi-GPU entry (both transports) against the single-GPU path, bit for bit
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29541 tools/dist_check.py 1024 all > gpurun_out/r02_dist_check_2gpu.log 2>&1; echo "dist_check exit $?"
tail -40 gpurun_out/r02_dist_check_2gpu.log
