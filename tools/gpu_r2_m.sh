#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_m.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu_m.log
tail -4 gpurun_out/r02_pytest_gpu_m.log
for L in ab/lib_head.so default; do
  if [ $L = default ]; then unset GEMMUL8_B200_LIB; else export GEMMUL8_B200_LIB=$PWD/$L; fi
  timeout 300 python tools/e2e_time.py 16384 14 5
done
unset GEMMUL8_B200_LIB
timeout 600 python tools/ab_time.py head=ab/lib_head.so,new=default "16384,16384,16384,16;16384,16384,16384,17;16384,16384,16384,18;16384,16384,16384,20" 1 > gpurun_out/r02_ab_bigroute.jsonl 2> gpurun_out/r02_ab_bigroute.err
python - <<PY
import json
for l in open("gpurun_out/r02_ab_bigroute.jsonl"):
    d=json.loads(l); print(d.get('shape'), d.get('variant'), d.get('us_best'), d.get('phases_us'), d.get('error'))
PY
