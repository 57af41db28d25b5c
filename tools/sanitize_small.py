"""Small invocations of every kernel family for compute-sanitizer (one tool per gpurun call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gemmul8_b200 as g

def real(m, n, k, N, fast, opA, opB, dt):
    rA, cA = (m, k) if opA == 0 else (k, m)
    rB, cB = (k, n) if opB == 0 else (n, k)
    A = g.phi_matrix(rA, cA, 0.5, dt); B = g.phi_matrix(rB, cB, 0.5, dt, seed=7)
    C = torch.zeros((n, m), dtype=dt, device="cuda")
    work = torch.zeros(g.workSize(m, n, k, N), dtype=torch.uint8, device="cuda")
    g.gemm(None, opA, opB, m, n, k, 1.0, A, rA, B, rB, 0.0, C, m, N, fast, work)
    torch.cuda.synchronize()

def cplx(m, n, k, N, fast, opA, opB, dt, ct):
    rA, cA = (m, k) if opA == 0 else (k, m)
    rB, cB = (k, n) if opB == 0 else (n, k)
    A = g.phi_matrix(rA, cA, 0.5, dt); B = g.phi_matrix(rB, cB, 0.5, dt, seed=7)
    C = torch.zeros((n, m), dtype=dt, device="cuda")
    work = torch.zeros(g.workSize(m, n, k, N, ct), dtype=torch.uint8, device="cuda")
    g.gemm(None, opA, opB, m, n, k, 1.0, A, rA, B, rB, 0.0, C, m, N, fast, work, computeType=ct)
    torch.cuda.synchronize()

real(300, 200, 260, 14, True, 0, 0, torch.float64)
real(129, 257, 131, 20, False, 1, 1, torch.float64)
real(77, 45, 33, 6, True, 0, 1, torch.float32)
for ct in (1, 2, 3):
    cplx(130, 70, 101, 9, True, 0, 2, torch.complex128, ct)
    cplx(70, 52, 100, 8, False, 1, 0, torch.complex64, ct)
g.set_option("gemm_pair", 1)
real(300, 520, 260, 14, True, 0, 0, torch.float64)
real(2048, 2304, 384, 14, True, 0, 0, torch.float64)          # >= 74 pairs' worth of items: placed path with the claim table, 256-bit stores
real(1040, 1100, 300, 9, True, 0, 0, torch.float64)           # ragged edges: predicated 4-byte stores
g.set_option("tma_store", 1)
real(2048, 2304, 384, 14, True, 0, 0, torch.float64)          # TMA-store epilogue, interior
real(1040, 1100, 300, 9, True, 0, 0, torch.float64)           # ... clipped at the edges of the matrix
g.set_option("tma_store", 0)
g.set_option("fused_k", 4096)
real(777, 1301, 300, 14, True, 0, 0, torch.float64)           # single-kernel product + CRT, ragged
real(300, 260, 200, 6, True, 0, 0, torch.float32)
g.set_option("fused_k", 0)
cplx(700, 600, 320, 14, True, 0, 0, torch.complex128, 3)      # Karatsuba: the combine passes of the pair kernel
print("sanitize workload done")
