#!/bin/bash
# ncu --set full of the single-kernel product + CRT at a short-k shape (and of the pair kernel + CRT kernel for comparison)
mkdir -p gpurun_out
python tools/profile_shape.py 16384 16384 512 14 2 64 > gpurun_out/plain_c.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:oz_gemm_crt -s 1 -c 1 -o gpurun_out/prof_r02_fused_k512 -f \
    python tools/profile_shape.py 16384 16384 512 14 2 64 > gpurun_out/ncu_c.log 2>&1
tail -3 gpurun_out/ncu_c.log
ls -la gpurun_out/prof_r02_fused_k512.ncu-rep
