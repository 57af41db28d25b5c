#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_v.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu_v.log
tail -5 gpurun_out/r02_pytest_gpu_v.log
for F in 1 0 1 0; do GEMMUL8_B200_SCALE_FORK=$F timeout 300 python tools/zgemm_time.py 8192 2>&1 | head -1; done | tee gpurun_out/r02_zgemm_overlap.jsonl
