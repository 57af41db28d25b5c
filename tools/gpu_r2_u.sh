#!/bin/bash
mkdir -p gpurun_out
python tools/sanitize_small.py > gpurun_out/r02_sanitize_plain.log 2>&1 && \
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_small.py > gpurun_out/r02_memcheck.log 2>&1; echo "memcheck exit $?"
tail -6 gpurun_out/r02_memcheck.log
