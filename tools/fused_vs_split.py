"""Single-kernel product + CRT (FLAG_FUSED_CRT) against product + residues + CRT kernel: same bits, and the time of both at
HPL-like trailing-update shapes (16384 x 16384 x k).  usage: fused_vs_split.py [moduli] [k,k,...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gemmul8_b200 as g

N = int(sys.argv[1]) if len(sys.argv) > 1 else 14
ks = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [256, 512, 1024, 2048, 4096]
m = n = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
g.init()
g.set_option("fused_k", 0)
for k in ks:
    A = g.phi_matrix(m, k, 0.5, torch.float64)
    B = g.phi_matrix(k, n, 0.5, torch.float64, seed=7)
    work = torch.empty(g.workSize(m, n, k, N), dtype=torch.uint8, device="cuda")
    C0 = torch.zeros((n, m), dtype=torch.float64, device="cuda")
    C1 = torch.zeros_like(C0)
    res = {"m": m, "n": n, "k": k, "moduli": N}
    for name, flags, C in (("split", 0, C0), ("fused", g.FLAG_FUSED_CRT, C1)):
        for _ in range(3):
            g.gemm(None, 0, 0, m, n, k, 1.0, A, m, B, k, 0.0, C, m, N, True, work, flags=flags)
        torch.cuda.synchronize()
        reps = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g.phase_log_collect()
        e0.record()
        for _ in range(reps):
            g.gemm(None, 0, 0, m, n, k, 1.0, A, m, B, k, 0.0, C, m, N, True, work, flags=flags | g.FLAG_PHASE_LOG)
        e1.record()
        torch.cuda.synchronize()
        ph, _ = g.phase_log_collect()
        ms = e0.elapsed_time(e1) / reps
        res[name] = {"us": round(ms * 1e3, 1), "TFLOPS": round(2.0 * m * n * k / ms / 1e9, 1),
                     "phases_us": [round(x / reps / 1e3, 1) for x in ph]}
    res["identical"] = bool(torch.equal(C0, C1)) and bool(C0.abs().sum() > 0)
    res["speedup"] = round(res["split"]["us"] / res["fused"]["us"], 3)
    print(json.dumps(res), flush=True)
    del A, B, work, C0, C1
