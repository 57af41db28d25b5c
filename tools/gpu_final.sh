#!/bin/bash
# Round-end GPU visit: parity tests, both bench arms, small sizes, per-kernel launch times at small sizes, ncu capture.
TAG=${1:-v9}
mkdir -p gpurun_out
bash tools/gpu_round.sh > /dev/null 2>&1
tail -2 gpurun_out/pytest_gpu.log
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python tools/small_sizes.py > gpurun_out/small_sizes_$TAG.jsonl 2>&1
for S in 1024 2048 4096; do
  ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active --clock-control none \
      -s 10 -c 6 --csv --log-file gpurun_out/launches_small_${S}_$TAG.csv python tools/profile_one_call.py $S 14 2 > /dev/null 2>&1
done
bash tools/gpu_profile.sh $TAG > gpurun_out/gpu_profile.log 2>&1
tail -2 gpurun_out/gpu_profile.log
python - <<PY
import json
d=json.load(open("gpurun_out/bench_ours.json")); r=json.load(open("gpurun_out/bench_ref.json"))
print("ours", round(d["value"],1), round(d["ms_per_step"],2), round(d["roofline"]["frac"],3), {k:(round(v,2) if isinstance(v,float) else v) for k,v in d["phases_ms"].items()}, "e2e", round(d["e2e"]["value"],1), round(d["e2e"]["ms_per_step"],1), "matched", d["accuracy_matched"]["moduli"], round(d["accuracy_matched"]["value"],1), d["clocks"])
print("ref", round(r["value"],1), round(r["ms_per_step"],1), round(r["device_resident"]["value"],1))
PY
cat gpurun_out/small_sizes_$TAG.jsonl
