import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gemmul8_b200 as g
m = n = k = 16384; N = 14
A = g.phi_matrix(m, k, 0.5, torch.float64); B = g.phi_matrix(k, n, 0.5, torch.float64)
work = torch.empty(g.workSize(m, n, k, N), dtype=torch.uint8, device="cuda")
C = torch.zeros((n, m), dtype=torch.float64, device="cuda")
def run(tag):
    for _ in range(2): g.gemm(None, 0, 0, m, n, k, 1.0, A, m, B, k, 0.0, C, m, N, True, work)
    torch.cuda.synchronize()
    t = [0.0]*4
    for _ in range(4):
        x = g.gemm(None, 0, 0, m, n, k, 1.0, A, m, B, k, 0.0, C, m, N, True, work, flags=g.FLAG_TIMERS)
        t = [a+b for a, b in zip(t, x)]
    print(tag, "gemm ms %.2f" % (t[1]/4/1e6), flush=True)
os.environ.pop("OZ_CLUSTER", None); os.environ.pop("OZ_MAP", None)
os.environ["OZ_GEMM_PAIR"] = "1"; os.environ["OZ_PAIR_MAP"] = "1"
for rep in range(3):
    for st in (4, 5, 6):
        os.environ["OZ_PAIR_STAGES"] = str(st); run("placed pair kernel, %d stages" % st)
