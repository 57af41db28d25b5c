#!/bin/bash
# Round-2 GPU visit D: full parity suite on the new pair-kernel epilogue, short-k timing, headline bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_d.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu_d.log
tail -6 gpurun_out/r02_pytest_gpu_d.log
timeout 600 python tools/fused_vs_split.py 14 256,512,1024,2048,4096 > gpurun_out/r02_fused_vs_split_d.jsonl 2> gpurun_out/r02_fused_vs_split_d.err; echo "exit $?"
cat gpurun_out/r02_fused_vs_split_d.jsonl; tail -3 gpurun_out/r02_fused_vs_split_d.err
timeout 600 python tools/small_sizes.py > gpurun_out/r02_small_sizes_d.jsonl 2>&1; cat gpurun_out/r02_small_sizes_d.jsonl
timeout 900 python bench.py --steps 10 --warmup 3 --no-config5 > gpurun_out/r02_bench_ours_d.json 2> gpurun_out/r02_bench_ours_d.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_ours_d.json"))
print("ours", round(d["value"],1), round(d["ms_per_step"],2), d["roofline"]["frac"], d["phases_ms"], "e2e", round(d["e2e"]["value"],1), d["accuracy_matched"]["moduli"], round(d["accuracy_matched"]["value"],1), d["clocks"])
PY
