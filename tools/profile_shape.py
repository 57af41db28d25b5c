"""A few gemm calls at one shape, for ncu.  usage: profile_shape.py m n k [moduli] [calls] [flags] [fused_k]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gemmul8_b200 as g

m, n, k = (int(x) for x in sys.argv[1:4])
N = int(sys.argv[4]) if len(sys.argv) > 4 else 14
calls = int(sys.argv[5]) if len(sys.argv) > 5 else 2
flags = int(sys.argv[6]) if len(sys.argv) > 6 else 0
if len(sys.argv) > 7:
    g.set_option("fused_k", int(sys.argv[7]))
g.init()
A = g.phi_matrix(m, k, 0.5, torch.float64)
B = g.phi_matrix(k, n, 0.5, torch.float64, seed=7)
work = torch.empty(g.workSize(m, n, k, N), dtype=torch.uint8, device="cuda")
C = torch.zeros((n, m), dtype=torch.float64, device="cuda")
for _ in range(calls):
    g.gemm(None, 0, 0, m, n, k, 1.0, A, m, B, k, 0.0, C, m, N, True, work, flags=flags)
torch.cuda.synchronize()
print("ok", float(C[5, 7]))
