#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/ab_time.py full=default,ldtm_only=ab/lib_epildtm.so,one_word_stores=ab/lib_oneword.so "16384,16384,512,14;16384,16384,2048,14" 1 > gpurun_out/r02_ab_mainloop2.jsonl 2> gpurun_out/r02_ab_mainloop2.err
tail -2 gpurun_out/r02_ab_mainloop2.err
python - <<PY
import json
for l in open("gpurun_out/r02_ab_mainloop2.jsonl"):
    d=json.loads(l); print(d.get('shape'), d.get('variant'), d.get('us_best'), d.get('phases_us'), d.get('error'))
PY
