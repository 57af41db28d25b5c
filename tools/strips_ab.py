import sys, os, json
sys.path.insert(0, "/root/repo")
import torch, gemmul8_b200 as g
S, N = 16384, 14
A = g.phi_matrix(S, S, 0.5, torch.float64); B = g.phi_matrix(S, S, 0.5, torch.float64, seed=7)
C = torch.zeros((S, S), dtype=torch.float64, device="cuda")
work = torch.empty(g.workSize(S, S, S, N), dtype=torch.uint8, device="cuda")
def run(flags, reps=10):
    for _ in range(3): g.gemm(None, 0, 0, S, S, S, 1.0, A, S, B, S, 0.0, C, S, N, True, work, flags=flags)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.gemm(None, 0, 0, S, S, S, 1.0, A, S, B, S, 0.0, C, S, N, True, work, flags=flags)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for i in range(3):
    print("default %.2f ms   strips %.2f ms" % (run(0), run(g.FLAG_STRIPS)), flush=True)
