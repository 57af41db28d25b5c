"""ZGEMM emulation at BASELINE config 4 (8192^3, 14 moduli): ms per call per compute type.  Honours GEMMUL8_B200_* options."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gemmul8_b200 as g
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
N = 14
g.init()
A = g.phi_matrix(S, S, 0.5, torch.complex128); B = g.phi_matrix(S, S, 0.5, torch.complex128, seed=7)
C = torch.zeros((S, S), dtype=torch.complex128, device="cuda")
for ct, name in ((3, "karatsuba"), (1, "bigmatrix"), (2, "classic")):
    work = torch.empty(g.workSize(S, S, S, N, ct), dtype=torch.uint8, device="cuda")
    for _ in range(2):
        g.gemm(None, 0, 0, S, S, S, 1.0, A, S, B, S, 0.0, C, S, N, True, work, computeType=ct)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8):
        g.gemm(None, 0, 0, S, S, S, 1.0, A, S, B, S, 0.0, C, S, N, True, work, computeType=ct)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 8
    print(json.dumps({"computeType": name, "fork": os.environ.get("GEMMUL8_B200_SCALE_FORK", "1"), "ms": round(ms, 3), "TFLOPS_8mnk": round(8.0 * S ** 3 / ms / 1e9, 1),
                      "checksum": float(torch.view_as_real(C).sum())}), flush=True)
    del work
