"""The synchronous (FLAG_TIMERS) flavour of a large call takes the column-strip pipeline: same bits as the asynchronous
call and as the phases in series, timers = exposed scaling / products / 0 / exposed CRT.  usage: strips_timers_check.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gemmul8_b200 as g
g.init()
m = n = k = 16384; N = 14
A = g.phi_matrix(m, k, 0.5, torch.float64); B = g.phi_matrix(k, n, 0.5, torch.float64, seed=7)
work = torch.empty(g.workSize(m, n, k, N), dtype=torch.uint8, device="cuda")
C0 = torch.zeros((n, m), dtype=torch.float64, device="cuda"); C1 = torch.zeros_like(C0); C2 = torch.zeros_like(C0)
c0 = g.get_option("strip_calls")
g.gemm(None, 0, 0, m, n, k, 1.0, A, m, B, k, 0.0, C0, m, N, True, work); torch.cuda.synchronize()
t = g.gemm(None, 0, 0, m, n, k, 1.0, A, m, B, k, 0.0, C1, m, N, True, work, flags=g.FLAG_TIMERS)
t0 = time.perf_counter()
t = g.gemm(None, 0, 0, m, n, k, 1.0, A, m, B, k, 0.0, C1, m, N, True, work, flags=g.FLAG_TIMERS)
wall = (time.perf_counter() - t0) * 1e3
calls = g.get_option("strip_calls") - c0
g.set_option("strips", 1)
ts = g.gemm(None, 0, 0, m, n, k, 1.0, A, m, B, k, 0.0, C2, m, N, True, work, flags=g.FLAG_TIMERS)
print(json.dumps({"strip_calls": calls, "timers_ms_pipeline": [x / 1e6 for x in t], "wall_ms_pipeline": wall, "timers_ms_series": [x / 1e6 for x in ts],
                  "equal_async": bool(torch.equal(C0, C1)), "equal_series": bool(torch.equal(C1, C2))}))
