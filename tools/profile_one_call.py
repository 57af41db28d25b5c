"""Counterpart of the reference's testing/profile_one_call.cu: a few gemmul8 calls at the benchmark
shape, for ncu.  usage: profile_one_call.py [size] [moduli] [calls] [flags]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gemmul8_b200 as g

S = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
N = int(sys.argv[2]) if len(sys.argv) > 2 else 14
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 2
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
m = n = k = S
A = g.phi_matrix(m, k, 0.5, torch.float64)
B = g.phi_matrix(k, n, 0.5, torch.float64)
work = torch.empty(g.workSize(m, n, k, N), dtype=torch.uint8, device="cuda")
C = torch.zeros((n, m), dtype=torch.float64, device="cuda")
for _ in range(calls):
    g.gemm(None, 0, 0, m, n, k, 1.0, A, m, B, k, 0.0, C, m, N, True, work, flags=flags)
torch.cuda.synchronize()
print("ok", float(C[5, 7]))
