"""Time the GEMM-phase variants on ONE box (clocks differ between boxes): prints ms per phase."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gemmul8_b200 as g

S = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
N = int(sys.argv[2]) if len(sys.argv) > 2 else 14
m = n = k = S
A = g.phi_matrix(m, k, 0.5, torch.float64)
B = g.phi_matrix(k, n, 0.5, torch.float64)
work = torch.empty(g.workSize(m, n, k, N), dtype=torch.uint8, device="cuda")
C = torch.zeros((n, m), dtype=torch.float64, device="cuda")

def run(name, flags, env, pair=False):
    os.environ["OZ_GEMM_PAIR"] = "" if pair else "0"
    if env is None:
        os.environ.pop("OZ_DEBUG_SCHED", None)
    else:
        os.environ["OZ_DEBUG_SCHED"] = env
    for _ in range(3):
        g.gemm(None, 0, 0, m, n, k, 1.0, A, m, B, k, 0.0, C, m, N, True, work, flags=flags)
    torch.cuda.synchronize()
    reps = 12
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.gemm(None, 0, 0, m, n, k, 1.0, A, m, B, k, 0.0, C, m, N, True, work, flags=flags)
    e1.record()
    torch.cuda.synchronize()
    ph = g.gemm(None, 0, 0, m, n, k, 1.0, A, m, B, k, 0.0, C, m, N, True, work, flags=flags | g.FLAG_TIMERS)   # (instrumented: in series)
    print(json.dumps({"variant": name, "total_ms": e0.elapsed_time(e1) / reps, "phases_ms_in_series": [p / 1e6 for p in ph]}), flush=True)

for rep in range(2):
    run("default: CTA-pair GEMM (cta_group::2, pairs placed like a plain launch), phases in series", 0, None, pair=True)
    run("single-CTA GEMM (cta_group::1), phases in series", 0, None)
    run("column-strip pipeline on 3 streams", g.FLAG_STRIPS, None)
    run("EXPERIMENT (wrong results): single-CTA, B tile loaded every other k-block only", 0, "halfb", pair=False)
    run("single-CTA, tile-major", 0, "tile", pair=False)
    run("fused", g.FLAG_FUSED_CRT, None)
    run("fused skipcrt", g.FLAG_FUSED_CRT, "skipcrt")
