#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python tools/sweep.py --out gpurun_out/r02_sweep_full.jsonl --csv gpurun_out/r02_sweep_full.csv > gpurun_out/r02_sweep.log 2>&1; echo "sweep exit $?"
python - <<PY
import json
for l in open("gpurun_out/r02_sweep_full.jsonl"):
    d=json.loads(l)
    if "error" in d: print(d); continue
    r=d.get("reference",{})
    print(d["types"], d["function"], d["computeType"], d["phi"], round(d["TFLOPS"],1), round(d["total_time_ms"],2), {k:round(v,2) for k,v in d["phases_ms"].items()}, "ref", round(r.get("TFLOPS",0),1), r.get("C_bit_identical"), "%.2e"%d["relerr_max"])
PY
