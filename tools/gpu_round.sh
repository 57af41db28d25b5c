#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_w.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu_w.log
tail -4 gpurun_out/r02_pytest_gpu_w.log
timeout 300 python tools/small_sizes.py --sizes 512,1024,2048,4096 2>&1 | tee gpurun_out/r02_small_sizes_w.jsonl
