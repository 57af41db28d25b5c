#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/ab_time.py base=default,hint1=ab/lib_hint1.so,hint2=ab/lib_hint2.so "16384,16384,16384,14;8192,8192,8192,14;16384,16384,2048,14;16384,16384,16384,8" 2 > gpurun_out/r02_ab_l2hint.jsonl 2> gpurun_out/r02_ab_l2hint.err
tail -2 gpurun_out/r02_ab_l2hint.err
python - <<PY
import json
for l in open("gpurun_out/r02_ab_l2hint.jsonl"):
    d=json.loads(l); print(d.get('shape'), d.get('variant'), d.get('us_best'), d.get('us_all'), d.get('phases_us'), d.get('error'))
PY
