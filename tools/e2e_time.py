"""End-to-end host-buffer call (gemmul8_b200_gemm_host) at the benchmark shape: ms per call, and C against the device-resident
call.  Honours GEMMUL8_B200_LIB (A/B of builds).  usage: e2e_time.py [size] [moduli] [reps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gemmul8_b200 as g

S = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
N = int(sys.argv[2]) if len(sys.argv) > 2 else 14
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
g.init()
m = n = k = S
A = g.phi_matrix(m, k, 0.5, torch.float64)
B = g.phi_matrix(k, n, 0.5, torch.float64, seed=7)
hA = torch.empty((k, m), dtype=torch.float64, pin_memory=True).copy_(A)
hB = torch.empty((n, k), dtype=torch.float64, pin_memory=True).copy_(B)
hC = torch.zeros((n, m), dtype=torch.float64, pin_memory=True)
scratch = torch.empty(g.host_scratch_size(0, 0, m, n, k, hA, m, hB, k, hC, m, N), dtype=torch.uint8, device="cuda")
for _ in range(2):
    g.gemm_host(0, 0, m, n, k, 1.0, hA, m, hB, k, 0.0, hC, m, N, True, scratch)
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.gemm_host(0, 0, m, n, k, 1.0, hA, m, hB, k, 0.0, hC, m, N, True, scratch)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
Cd = torch.zeros((n, m), dtype=torch.float64, device="cuda")
work = torch.empty(g.workSize(m, n, k, N), dtype=torch.uint8, device="cuda")
g.gemm(None, 0, 0, m, n, k, 1.0, A, m, B, k, 0.0, Cd, m, N, True, work)
torch.cuda.synchronize()
print(json.dumps({"lib": os.environ.get("GEMMUL8_B200_LIB", "default"), "ms_best": min(ts), "ms_all": [round(t, 2) for t in ts],
                  "TFLOPS": 2.0 * m * n * k / min(ts) / 1e9, "equals_device_call": bool(torch.equal(hC, Cd.cpu()))}))
