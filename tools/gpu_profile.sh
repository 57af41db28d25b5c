#!/bin/bash
# ncu evidence for one gemm call at the benchmark shape (B200_PROFILING.md recipe): run plain first, then the launch
# list, then one --set full capture (all kernels of the second call).  Outputs under gpurun_out/.
set -x
TAG=${1:-v6}
mkdir -p gpurun_out
python tools/profile_one_call.py 16384 14 2 > gpurun_out/plain_one_$TAG.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$TAG.csv \
    python tools/profile_one_call.py 16384 14 2 > gpurun_out/ncu_l_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -s 10 -c 6 -o gpurun_out/prof_r01_$TAG -f \
    python tools/profile_one_call.py 16384 14 2 > gpurun_out/ncu_f_$TAG.log 2>&1
ncu -i gpurun_out/prof_r01_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_full_${TAG}_raw.csv 2> /dev/null
tail -3 gpurun_out/ncu_f_$TAG.log
