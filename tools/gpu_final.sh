#!/bin/bash
# Round-end GPU visit on one B200: parity suite, smoke, both bench arms.  Outputs under gpurun_out/.
TAG=${1:-final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "reference arm exit $?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ours_$TAG.json 2> gpurun_out/bench_ours_$TAG.err; echo "our arm exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_ours_$TAG.json")); r=json.load(open("gpurun_out/bench_ref_$TAG.json"))
print("ours", round(d["value"],1), round(d["ms_per_step"],2), "frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"],1), d["clocks"])
print("ref ", round(r["value"],1), round(r["ms_per_step"],2), "e2e", round(r["e2e"]["value"],1), r["clocks"])
PY
