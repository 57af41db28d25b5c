#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29541 tools/dist_check.py 1024 copy > gpurun_out/r02_dist_check_${N}gpu_v3.log 2>&1; echo "dist_check exit $?"
grep -E "DIST|rror" gpurun_out/r02_dist_check_${N}gpu_v3.log | tail -3
run() {  # name, args
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 10 --warmup 3 --exchange copy "$@" > gpurun_out/r02_mp_${N}gpu_$name.json 2> gpurun_out/r02_mp_${N}gpu_$name.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02_mp_${N}gpu_$name.json") if l.startswith("{")][-1])
    print("$name", round(d["value"],1), round(d["ms_per_step"],2), {k:(round(v,2) if isinstance(v,float) else v) for k,v in d["phases_ms"].items() if k!='note'}, d.get("parity",{}).get("ok"), d.get("config5"))
except Exception as e: print("$name no json", e)
PY
}
run final
