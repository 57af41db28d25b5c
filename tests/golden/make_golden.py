#!/usr/bin/env python3
"""Generate golden vectors from the UNMODIFIED reference (oracle/_ref/libgemmul8_ref.so) on a GPU box.

    gpurun -- python tests/golden/make_golden.py        # writes gpurun_out/golden/*.npz
    cp gpurun_out/golden/*.npz tests/golden/

Each file holds the reference's inputs (its own phi-matrix generator, reproduced by
gemmul8_aux_phi_matrix) and everything gemmul8::gemm leaves behind: shifts, int8 slices, per-modulus
residues and C.  The CPU oracle and the CUDA path are both checked against these (tests/test_golden.py).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import gemmul8_b200 as g
import oracle

TORCH = {"f32": torch.float32, "f64": torch.float64}
CASES = [
    # name, m, n, k, N, fast, dtA, dtB, dtC, opA, opB, alpha, beta, phi
    ("d_fast_n14", 96, 80, 112, 14, 1, "f64", "f64", "f64", 0, 0, 1.0, 0.0, 0.5),
    ("d_accu_n14", 96, 80, 112, 14, 0, "f64", "f64", "f64", 0, 0, 1.0, 0.0, 0.5),
    ("d_fast_n8_tt", 70, 52, 100, 8, 1, "f64", "f64", "f64", 1, 1, 1.0, 0.0, 1.0),
    ("d_fast_n20_ragged", 45, 33, 77, 20, 1, "f64", "f64", "f64", 0, 1, 1.0, 0.0, 2.0),
    ("d_fast_n2", 32, 32, 32, 2, 1, "f64", "f64", "f64", 0, 0, 1.0, 0.0, 0.5),
    ("d_fast_n7_ab", 64, 48, 64, 7, 1, "f64", "f64", "f64", 0, 0, 0.75, -1.5, 0.5),
    ("d_fast_n14_ab", 64, 48, 64, 14, 1, "f64", "f64", "f64", 1, 0, 0.75, -1.5, 0.5),
    ("s_fast_n6", 96, 80, 112, 6, 1, "f32", "f32", "f32", 0, 0, 1.0, 0.0, 0.5),
    ("s_accu_n8", 60, 72, 90, 8, 0, "f32", "f32", "f32", 0, 0, 1.0, 0.0, 1.0),
    ("dfd_fast_n12", 80, 64, 96, 12, 1, "f64", "f32", "f64", 0, 0, 1.0, 0.0, 0.5),
    ("fdd_accu_n10", 80, 64, 96, 10, 0, "f32", "f64", "f64", 0, 0, 1.0, 0.0, 0.5),
    ("dff_fast_n6", 80, 64, 96, 6, 1, "f64", "f32", "f32", 0, 0, 1.0, 0.0, 0.5),
    ("fdf_fast_n6", 80, 64, 96, 6, 1, "f32", "f64", "f32", 1, 1, 1.0, 0.0, 0.5),
]


CTORCH = {"c32": torch.complex64, "c64": torch.complex128}
# complex cases: name, m, n, k, N, fast, dtA, dtB, dtC, opA, opB, computeType, phi   (alpha = 1, beta = 0: the only
# setting every reference kernel honours; big-matrix needs k % 4 == 0 in the reference)
COMPLEX_CASES = [
    ("z_big_fast_n14", 48, 40, 64, 14, 1, "c64", "c64", "c64", 0, 0, 1, 0.5),
    ("z_kara_fast_n14", 48, 40, 60, 14, 1, "c64", "c64", "c64", 0, 0, 3, 0.5),
    ("z_classic_fast_n9_ct", 40, 36, 52, 9, 1, "c64", "c64", "c64", 2, 1, 2, 1.0),
    ("c_kara_accu_n15", 48, 40, 64, 15, 0, "c32", "c32", "c32", 0, 0, 3, 0.5),      # one_accuracy_complex.cu's setting
    ("c_big_fast_n6_tc", 36, 44, 52, 6, 1, "c32", "c32", "c32", 1, 2, 1, 0.5),
    ("z_big_accu_n8", 40, 40, 48, 8, 0, "c64", "c64", "c64", 0, 2, 1, 0.5),
    ("zcz_kara_fast_n12", 40, 32, 48, 12, 1, "c64", "c32", "c64", 0, 0, 3, 0.5),
    ("czc_classic_fast_n6", 40, 32, 47, 6, 1, "c32", "c64", "c32", 0, 0, 2, 0.5),
]


def complex_cases(out):
    for (name, m, n, k, N, fast, dtA, dtB, dtC, opA, opB, ct, phi) in COMPLEX_CASES:
        rA, cA = (m, k) if opA == 0 else (k, m)
        rB, cB = (k, n) if opB == 0 else (n, k)
        A = g.phi_matrix(rA, cA, phi, CTORCH[dtA], seed=123456)
        B = g.phi_matrix(rB, cB, phi, CTORCH[dtB], seed=654321)
        Cr = torch.zeros((n, m), dtype=CTORCH[dtC], device="cuda")
        ws = oracle.ref_worksize(m, n, k, N, ct)
        work = torch.zeros(ws, dtype=torch.uint8, device="cuda")
        oracle.ref_gemm(opA, opB, m, n, k, 1.0, A, rA, B, rB, 0.0, Cr, m, N, fast, work, ct)
        ref_writes_c = ct == 1 or N <= 7 or dtC == "c32"
        if not ref_writes_c:   # reference defect: C untouched -> take C from its big-matrix mode (same residues)
            assert (Cr == 0).all()
            wb = torch.zeros(oracle.ref_worksize(m, n, k, N, 1), dtype=torch.uint8, device="cuda")
            oracle.ref_gemm(opA, opB, m, n, k, 1.0, A, rA, B, rB, 0.0, Cr, m, N, fast, wb, 1)
        L = g.work_layout(m, n, k, N, ct)
        assert L.total == ws
        v = g.work_views_complex(work, L, N, m, n, k, ct)
        arrays = {key: (val[:, :2 * m] if key == "A8i" else val[:, :m] if key.startswith("A8i_") else val).cpu().numpy()
                  for key, val in v.items() if key != "C8u"}
        np.savez_compressed(os.path.join(out, name + ".npz"),
                            meta=np.array([m, n, k, N, fast, opA, opB, ct], np.int64), phi=phi, dtypes=np.array([dtA, dtB, dtC]),
                            c_from_big_matrix_mode=np.array(int(not ref_writes_c)),
                            A=A.cpu().numpy(), B=B.cpu().numpy(), C=Cr.cpu().numpy(), **arrays)
        print("wrote", name, flush=True)


def main():
    out = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out, exist_ok=True)
    if "--complex-only" in sys.argv:
        return complex_cases(out)
    for (name, m, n, k, N, fast, dtA, dtB, dtC, opA, opB, alpha, beta, phi) in CASES:
        rA, cA = (m, k) if opA == 0 else (k, m)
        rB, cB = (k, n) if opB == 0 else (n, k)
        A = g.phi_matrix(rA, cA, phi, TORCH[dtA], seed=123456)
        B = g.phi_matrix(rB, cB, phi, TORCH[dtB], seed=654321)
        C0 = g.phi_matrix(m, n, phi, TORCH[dtC], seed=777)
        Cr = C0.clone()
        ws = oracle.ref_worksize(m, n, k, N)
        work = torch.zeros(ws, dtype=torch.uint8, device="cuda")
        oracle.ref_gemm(opA, opB, m, n, k, alpha, A, rA, B, rB, beta, Cr, m, N, fast, work)
        L = g.work_layout(m, n, k, N)
        assert L.total == ws
        v = g.work_views(work, L, N, m, n)
        np.savez_compressed(os.path.join(out, name + ".npz"),
                            meta=np.array([m, n, k, N, fast, opA, opB], np.int64), alpha=alpha, beta=beta, phi=phi,
                            dtypes=np.array([dtA, dtB, dtC]),
                            A=A.cpu().numpy(), B=B.cpu().numpy(), C0=C0.cpu().numpy(), C=Cr.cpu().numpy(),
                            sftA=v["sftA"].cpu().numpy(), sftB=v["sftB"].cpu().numpy(),
                            A8i=v["A8i"][:, :m].cpu().numpy(), B8i=v["B8i"].cpu().numpy(),
                            C8u=v["C8u"][:, :, :m].cpu().numpy())
        print("wrote", name, flush=True)
    complex_cases(out)


if __name__ == "__main__":
    main()
