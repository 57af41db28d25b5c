#!/bin/bash
mkdir -p gpurun_out
python tools/profile_shape.py 16384 16384 512 14 2 0 > gpurun_out/plain_e.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:oz_gemm_pair -s 1 -c 1 -o gpurun_out/prof_r02_pair_k512 -f \
    python tools/profile_shape.py 16384 16384 512 14 2 0 > gpurun_out/ncu_e.log 2>&1
tail -2 gpurun_out/ncu_e.log
