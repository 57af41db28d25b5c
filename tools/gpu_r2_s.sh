#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/ab_time.py tma=default,hint=ab/lib_sthint.so "16384,16384,512,14;16384,16384,1024,14;16384,16384,2048,14;16384,16384,16384,14" 2 > gpurun_out/r02_ab_sthint.jsonl 2> gpurun_out/r02_ab_sthint.err
tail -2 gpurun_out/r02_ab_sthint.err
python - <<PY
import json
for l in open("gpurun_out/r02_ab_sthint.jsonl"):
    d=json.loads(l); print(d.get('shape'), d.get('variant'), d.get('us_best'), d.get('us_all'), d.get('phases_us'), d.get('sum'), d.get('error'))
PY
