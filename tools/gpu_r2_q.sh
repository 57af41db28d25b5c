#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_q.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu_q.log
tail -5 gpurun_out/r02_pytest_gpu_q.log
timeout 300 python tools/ab_time.py fork=default,nofork=default:GEMMUL8_B200_SCALE_FORK=0 "1024,1024,1024,14;2048,2048,2048,14;1536,1536,1536,14;512,512,512,14" 2 > gpurun_out/r02_ab_fork.jsonl 2> gpurun_out/r02_ab_fork.err
python - <<PY
import json
for l in open("gpurun_out/r02_ab_fork.jsonl"):
    d=json.loads(l); print(d.get('shape'), d.get('variant'), d.get('us_best'), d.get('us_all'), d.get('phases_us'), d.get('error'))
PY
python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02_bench_ref_q.json 2> gpurun_out/r02_bench_ref_q.err; echo "ref exit $?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_ours_q.json 2> gpurun_out/r02_bench_ours_q.err; echo "ours exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_ours_q.json")); r=json.load(open("gpurun_out/r02_bench_ref_q.json"))
print("ours", round(d["value"],1), round(d["ms_per_step"],2), "frac", round(d["roofline"]["frac"],3), {k:(round(v,2) if isinstance(v,float) else v) for k,v in d["phases_ms"].items()}, "e2e", round(d["e2e"]["value"],1), round(d["e2e"]["ms_per_step"],1), "matched", d["accuracy_matched"]["moduli"], round(d["accuracy_matched"]["value"],1), d["clocks"], d["config5"].get("ms_per_step"), d["roofline"]["int8_peak_measured"])
print("ref", round(r["value"],1), round(r["ms_per_step"],1), "e2e", round(r["e2e"]["value"],1), r["clocks"])
PY
