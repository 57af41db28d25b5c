#!/bin/bash
# ncu evidence for the column-strip pipeline (the default of large calls): launch list of the bench command and one
# --set full capture of the four product launches of one call.  Outputs under gpurun_out/.
TAG=${1:-r02_strips}
mkdir -p gpurun_out
timeout 120 python bench.py --steps 2 --warmup 1 --no-e2e --no-config5 --no-cpu-baseline > gpurun_out/plain_bench_$TAG.log 2>&1 || { echo "plain bench failed"; exit 1; }
timeout 170 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/${TAG}_ncu_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-config5 --no-cpu-baseline > gpurun_out/ncu_l_$TAG.log 2>&1
echo "launch list exit $?"; grep -c oz_gemm_pair gpurun_out/${TAG}_ncu_launches_bench.csv
timeout 60 python tools/profile_one_call.py 16384 14 2 > gpurun_out/plain_one_$TAG.log 2>&1 || { echo "plain call failed"; exit 1; }
timeout 200 ncu --set full --clock-control none -k regex:"oz_gemm_pair" -s 4 -c 4 -o gpurun_out/prof_${TAG}_16384 -f \
    python tools/profile_one_call.py 16384 14 2 > gpurun_out/ncu_f_$TAG.log 2>&1
echo "full capture exit $?"; tail -2 gpurun_out/ncu_f_$TAG.log
ncu -i gpurun_out/prof_${TAG}_16384.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_full_raw.csv 2> /dev/null
ls -la gpurun_out/prof_${TAG}_16384.ncu-rep gpurun_out/${TAG}_ncu_full_raw.csv gpurun_out/${TAG}_ncu_launches_bench.csv
