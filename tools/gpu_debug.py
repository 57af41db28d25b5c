#!/usr/bin/env python3
"""First-contact diagnostics on a B200: every stage of our path against the unmodified reference
(oracle/_ref/libgemmul8_ref.so) and the CUDA-core cross-check.  Prints, never asserts."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import gemmul8_b200 as g
import oracle


def run_case(m, n, k, N, fast, dtA=torch.float64, dtB=torch.float64, dtC=torch.float64, opA=0, opB=0, phi=0.5, simt=False, ref=True):
    dev = "cuda"
    rA, cA = (m, k) if opA == 0 else (k, m)
    rB, cB = (k, n) if opB == 0 else (n, k)
    A = g.phi_matrix(rA, cA, phi, dtA)
    B = g.phi_matrix(rB, cB, phi, dtB)
    ws = g.workSize(m, n, k, N)
    L = g.work_layout(m, n, k, N)
    work = torch.zeros(ws, dtype=torch.uint8, device=dev)
    Cm = torch.zeros((n, m), dtype=dtC, device=dev)
    flags = g.FLAG_GEMM_SIMT if simt else 0
    t0 = time.time()
    g.gemm(None, opA, opB, m, n, k, 1.0, A, rA, B, rB, 0.0, Cm, m, N, fast, work, flags=flags)
    torch.cuda.synchronize()
    t1 = time.time() - t0
    v = g.work_views(work, L, N, m, n)
    tag = f"m={m} n={n} k={k} N={N} fast={int(fast)} op=({opA},{opB}) {str(dtA)[6:]}x{str(dtB)[6:]}->{str(dtC)[6:]} simt={int(simt)}"
    msg = [tag, f"t={t1*1e3:.1f}ms"]
    if ref and oracle.have_ref():
        rws = oracle.ref_worksize(m, n, k, N)
        assert rws == ws, (rws, ws)
        rwork = torch.zeros(rws, dtype=torch.uint8, device=dev)
        Cr = torch.zeros((n, m), dtype=dtC, device=dev)
        oracle.ref_gemm(opA, opB, m, n, k, 1.0, A, rA, B, rB, 0.0, Cr, m, N, fast, rwork)
        rv = g.work_views(rwork, L, N, m, n)
        for name in ("sftA", "sftB"):
            d = (v[name] != rv[name]).sum().item()
            msg.append(f"{name} diff={d}")
        for name in ("A8i", "B8i", "C8u"):
            a, b = v[name], rv[name]
            if name == "A8i":
                a, b = a[:, :m], b[:, :m]
            if name == "C8u":
                a, b = a[:, :, :m], b[:, :, :m]
            d = (a != b).sum().item()
            msg.append(f"{name} diff={d}/{a.numel()}")
        dC = (Cm != Cr).sum().item()
        rel = ((Cm.double() - Cr.double()).abs() / Cr.double().abs().clamp_min(1e-300)).max().item()
        msg.append(f"C diff={dC}/{Cm.numel()} maxrel={rel:.2e}")
    if dtA == torch.float64 and dtB == torch.float64 and m * n * k <= 2048 ** 3:
        C1, C2 = g.dd_gemm(m, n, k, A, rA, B, rB, opA != 0, opB != 0)
        err = ((Cm.double() - C1 - C2) / C1).abs()
        msg.append(f"relerr max={err.max().item():.6e} med={err.median().item():.3e}")
    print(" | ".join(msg), flush=True)


def main():
    print(torch.cuda.get_device_name(0), g.version(), "ref:", oracle.have_ref(), flush=True)
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "simt"):
        run_case(128, 256, 128, 14, True, simt=True)
        run_case(100, 70, 200, 14, True, simt=True)
        run_case(100, 70, 200, 8, False, simt=True)
        run_case(64, 64, 64, 6, True, dtA=torch.float32, dtB=torch.float32, dtC=torch.float32, simt=True)
    if which in ("all", "tc"):
        run_case(128, 256, 128, 2, True)
        run_case(128, 256, 512, 14, True)
        run_case(100, 70, 200, 14, True)
        run_case(300, 520, 1000, 14, True)
        run_case(1024, 1024, 1024, 14, True)
        run_case(1024, 1024, 1024, 14, False)
        run_case(1024, 1024, 1024, 14, True, opA=1, opB=1)
        run_case(1000, 900, 1100, 6, True, dtA=torch.float32, dtB=torch.float32, dtC=torch.float32)
        run_case(1000, 900, 1100, 12, True, dtA=torch.float64, dtB=torch.float32, dtC=torch.float64)
        run_case(4096, 4096, 4096, 14, True)


if __name__ == "__main__":
    main()
