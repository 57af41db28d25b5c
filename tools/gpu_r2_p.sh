#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python tools/ref_drivers.py run --out gpurun_out/refdrivers_r02_flops --drivers test_mixed_double,test_float_complex --checks flops_check --timeout 800 > gpurun_out/r02_refdrivers_flops_run.log 2>&1
tail -6 gpurun_out/r02_refdrivers_flops_run.log
