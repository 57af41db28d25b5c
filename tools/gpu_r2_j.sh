#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29541 tools/dist_check.py 1024 copy > gpurun_out/r02_dist_check_${N}gpu_v2.log 2>&1; echo "dist_check exit $?"
grep -E "DIST|exchange|unavailable|rror" gpurun_out/r02_dist_check_${N}gpu_v2.log | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 10 --warmup 3 --exchange copy $2 > gpurun_out/r02_bench_${N}gpu_copy_v2.json 2> gpurun_out/r02_bench_${N}gpu_copy_v2.err; echo "bench exit $?"
tail -2 gpurun_out/r02_bench_${N}gpu_copy_v2.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02_bench_${N}gpu_copy_v2.json") if l.startswith("{")][-1])
    print(round(d["value"],1), round(d["ms_per_step"],2), d["phases_ms"], d.get("parity",{}).get("ok"), d.get("config5"))
except Exception as e: print("no json", e)
PY
