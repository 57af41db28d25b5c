#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_complex_gpu.py -m gpu -x -q > gpurun_out/r02_pytest_g.log 2>&1; tail -3 gpurun_out/r02_pytest_g.log
timeout 900 python tools/ab_time.py st256=default,st128=ab/lib_st128.so "16384,16384,512,14;16384,16384,1024,14;16384,16384,2048,14;16384,16384,4096,14;16384,16384,16384,14;8192,8192,8192,14" 2 > gpurun_out/r02_ab_g4.jsonl 2> gpurun_out/r02_ab_g4.err
tail -3 gpurun_out/r02_ab_g4.err
python - <<PY
import json
for l in open("gpurun_out/r02_ab_g4.jsonl"):
    d=json.loads(l); print(d.get('shape'), d.get('variant'), d.get('us_best'), d.get('us_all'), d.get('phases_us'), d.get('error'))
PY
