#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/ab_time.py base=default,strips=default:GEMMUL8_B200_STRIP_K=8192 "16384,16384,256,14;16384,16384,512,14;16384,16384,1024,14;16384,16384,2048,14;16384,16384,4096,14;8192,8192,1024,14;32768,16384,1024,14" 2 > gpurun_out/r02_ab_strips.jsonl 2> gpurun_out/r02_ab_strips.err
tail -2 gpurun_out/r02_ab_strips.err
python - <<PY
import json
for l in open("gpurun_out/r02_ab_strips.jsonl"):
    d=json.loads(l); print(d.get('shape'), d.get('variant'), d.get('us_best'), d.get('us_all'), d.get('phases_us'), d.get('sum'), d.get('error'))
PY
