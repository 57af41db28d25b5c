#!/bin/bash
# Round-2 GPU visit A: parity suite (incl. the BASELINE-size bit-compares and the runtime tests), both bench arms,
# the accuracy checks of the three reference drivers that round 1 never ran.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_a.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu_a.log
tail -15 gpurun_out/r02_pytest_gpu_a.log
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02_bench_ref_a.json 2> gpurun_out/r02_bench_ref_a.err; echo "ref bench exit $?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_ours_a.json 2> gpurun_out/r02_bench_ours_a.err; echo "bench exit $?"
tail -3 gpurun_out/r02_bench_ours_a.err
cat gpurun_out/r02_bench_ref_a.json gpurun_out/r02_bench_ours_a.json
timeout 900 python tools/ref_drivers.py run --out gpurun_out/refdrivers_r02 --drivers test_mixed_double,test_mixed_float,test_float_complex --checks accuracy_check > gpurun_out/r02_refdrivers_run.log 2>&1
tail -8 gpurun_out/r02_refdrivers_run.log
