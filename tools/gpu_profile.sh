#!/bin/bash
# ncu evidence for one gemm call at the benchmark shape (B200_PROFILING.md recipe): run plain first, then the launch
# list of the bench command, then one --set full capture of the six kernels of the second call.  Outputs under gpurun_out/.
TAG=${1:-r02}
mkdir -p gpurun_out
python tools/profile_one_call.py 16384 14 2 > gpurun_out/plain_one_$TAG.log 2>&1 || exit 1
python bench.py --steps 2 --warmup 1 --no-e2e --no-config5 --no-cpu-baseline > gpurun_out/plain_bench_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_ncu_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-config5 --no-cpu-baseline > gpurun_out/ncu_l_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"oz_gemm|crt_kernel|encode_|fast_shift" -s 6 -c 6 -o gpurun_out/prof_${TAG}_16384 -f \
    python tools/profile_one_call.py 16384 14 2 > gpurun_out/ncu_f_$TAG.log 2>&1
tail -2 gpurun_out/ncu_f_$TAG.log
ls -la gpurun_out/prof_${TAG}_16384.ncu-rep gpurun_out/${TAG}_ncu_launches_bench.csv
