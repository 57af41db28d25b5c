#!/bin/bash
# Round-2 GPU visit B: the single-kernel product + CRT: parity tests, then time against the two-kernel path.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "fused or no_writes" > gpurun_out/r02_pytest_fused.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_fused.log
tail -12 gpurun_out/r02_pytest_fused.log
timeout 600 python tools/fused_vs_split.py 14 256,512,1024,2048,4096,16384 > gpurun_out/r02_fused_vs_split.jsonl 2> gpurun_out/r02_fused_vs_split.err; echo "exit $?"
cat gpurun_out/r02_fused_vs_split.jsonl; tail -3 gpurun_out/r02_fused_vs_split.err
