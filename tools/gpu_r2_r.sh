#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/ab_time.py st6=default,st7=default:OZ_PAIR_STAGES=7,st6b=default,st7b=default:OZ_PAIR_STAGES=7 "16384,16384,16384,14;8192,8192,8192,14;16384,16384,4096,14" 2 > gpurun_out/r02_ab_stages7.jsonl 2> gpurun_out/r02_ab_stages7.err
python - <<PY
import json
for l in open("gpurun_out/r02_ab_stages7.jsonl"):
    d=json.loads(l); print(d.get('shape'), d.get('variant'), d.get('us_best'), d.get('us_all'), d.get('phases_us'), d.get('error'))
PY
