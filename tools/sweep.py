#!/usr/bin/env python3
"""The reference's benchmark sweep (GEMMul8/testing/test_double.cu, test_float.cu, test_mixed_double.cu,
test_float_complex.cu: time_check + accuracy columns) on this library, the unmodified reference
library and native cuBLAS, one JSON line per configuration (BASELINE.json configs 2-4).

    python tools/sweep.py [--quick] [--out gpurun_out/sweep.jsonl]

Columns follow the reference's CSVs (oz2_results_*_time_*.csv): phi, function, relerr_max, relerr_med,
TFLOPS, total_time and the four phase times.  `function` is OS2-fast-N / OS2-accu-N as in the reference.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import gemmul8_b200 as g
import oracle

DT = {"d": torch.float64, "s": torch.float32, "z": torch.complex128, "c": torch.complex64}
CT_NAME = {0: "real", 1: "bigmatrix", 2: "classic", 3: "karatsuba"}


def timed(fn, warm, reps):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def truth_sample(A, B, m, n, k, ns=48):
    """rows / cols sample of C and its truth: double-double for real fp64 inputs, fp64 (complex128) product of the
    widened inputs otherwise (the reference's one_accuracy.cu does the same for float inputs)."""
    gen = torch.Generator().manual_seed(7)
    rows = torch.randperm(m, generator=gen)[:min(ns, m)].sort().values.cuda()
    cols = torch.randperm(n, generator=gen)[:min(ns, n)].sort().values.cuda()
    if A.dtype == torch.float64 and B.dtype == torch.float64:
        C1, C2 = g.dd_gemm(m, n, k, A, m, B, k, rows=rows.to(torch.int32).contiguous(), cols=cols.to(torch.int32).contiguous())
        return rows, cols, C1, C2
    wide = torch.complex128 if A.is_complex() else torch.float64
    T = B[cols].to(wide) @ A[:, rows].to(wide)
    return rows, cols, T, torch.zeros_like(T)


def relerr(C, rows, cols, C1, C2):
    sub = C[cols][:, rows].to(C1.dtype)
    err = ((sub - C1 - C2).abs() / C1.abs())
    return err.max().item(), err.median().item()


def one(kindA, kindB, kindC, size, N, fast, phi, ct, reps, with_ref, with_native=True):
    m = n = k = size
    A = g.phi_matrix(m, k, phi, DT[kindA], seed=123456)
    B = g.phi_matrix(k, n, phi, DT[kindB], seed=123456)
    C = torch.zeros((n, m), dtype=DT[kindC], device="cuda")
    ws = g.workSize(m, n, k, N, ct)
    work = torch.empty(ws, dtype=torch.uint8, device="cuda")
    cplx = kindC in "zc"
    flops = (8.0 if cplx else 2.0) * m * n * k
    row = {"types": kindA + kindB + kindC, "m": m, "n": n, "k": k, "phi": phi, "num_moduli": N,
           "function": f"OS2-{'fast' if fast else 'accu'}-{N}", "computeType": CT_NAME[ct], "workspace_GiB": ws / 2 ** 30}

    ph = [0.0] * 4

    def ours():
        t = g.gemm(None, 0, 0, m, n, k, 1.0, A, m, B, k, 0.0, C, m, N, fast, work, computeType=ct, flags=g.FLAG_TIMERS)
        for i in range(4):
            ph[i] += t[i]

    ms = timed(lambda: g.gemm(None, 0, 0, m, n, k, 1.0, A, m, B, k, 0.0, C, m, N, fast, work, computeType=ct), 2, reps)
    ours()
    rows, cols, C1, C2 = truth_sample(A, B, m, n, k)
    emax, emed = relerr(C, rows, cols, C1, C2)
    row.update({"TFLOPS": flops / ms / 1e9, "total_time_ms": ms, "relerr_max": emax, "relerr_med": emed,
                "phases_ms": {"scaling": ph[0] / 1e6, "int8_gemm_residues": ph[1] / 1e6, "crt": ph[3] / 1e6}})
    if with_native:
        Cn = torch.empty((n, m), dtype=torch.promote_types(A.dtype, B.dtype), device="cuda")
        An, Bn = A.to(Cn.dtype), B.to(Cn.dtype)
        msn = timed(lambda: torch.matmul(Bn, An, out=Cn), 1, max(2, reps // 2))
        nmax, nmed = relerr(Cn, rows, cols, C1, C2)
        row["native"] = {"TFLOPS": flops / msn / 1e9, "relerr_max": nmax, "relerr_med": nmed}
        del Cn, An, Bn
    if with_ref and oracle.have_ref():
        try:
            Cr = torch.zeros_like(C)
            rph = [0.0] * 4

            def ref():
                t = oracle.ref_gemm(0, 0, m, n, k, 1.0, A, m, B, k, 0.0, Cr, m, N, fast, work, ct)
                for i in range(4):
                    rph[i] = t[i]

            msr = timed(ref, 1, max(2, reps // 2))
            rmax, rmed = relerr(Cr, rows, cols, C1, C2)
            row["reference"] = {"TFLOPS": flops / msr / 1e9, "total_time_ms": msr, "relerr_max": rmax, "relerr_med": rmed,
                                "phases_ms": [x / 1e6 for x in rph], "C_bit_identical": bool(torch.equal(torch.view_as_real(C) if cplx else C, torch.view_as_real(Cr) if cplx else Cr))}
            del Cr
        except Exception as e:   # the reference faults on some shapes (e.g. big-matrix with k % 4 != 0)
            row["reference"] = {"error": str(e)[:200]}
    return row


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.jsonl"))
    ap.add_argument("--size", type=int, default=16384)
    ap.add_argument("--csize", type=int, default=8192)
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--csv", default=None, help="also write the rows in the reference's CSV format "
                    "(oz2_results_*_time_*.csv: GEMMul8/testing/test_double.cu:204-213), one file, ours / reference / native rows")
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    S, CS = a.size, a.csize
    reps = 3 if a.quick else 5
    cfgs = []
    # config 2: DGEMM emulation, moduli sweep, fast + accurate, phi in {0.5, 1, 2}
    mods = [8, 14, 20] if a.quick else list(range(8, 21))
    for N in mods:
        for fast in (True, False):
            cfgs.append(("d", "d", "d", S, N, fast, 0.5, 0))
    for phi in (1.0, 2.0):
        for fast in (True, False):
            cfgs.append(("d", "d", "d", S, 14, fast, phi, 0))
    # config 3: SGEMM emulation (6 moduli; the reference's figure also shows 7, 8) and FP64 x FP32 -> FP64
    for N in ([6] if a.quick else [6, 7, 8]):
        cfgs.append(("s", "s", "s", S, N, True, 0.5, 0))
    cfgs.append(("s", "s", "s", S, 6, False, 0.5, 0))
    cfgs.append(("d", "s", "d", S, 14, True, 0.5, 0))
    cfgs.append(("d", "s", "s", S, 6, True, 0.5, 0))
    # config 4: ZGEMM emulation 8192^3, Karatsuba vs big matrix (and classic), 14 moduli; CGEMM as test_float_complex.cu
    for ct in (3, 1, 2):
        cfgs.append(("z", "z", "z", CS, 14, True, 0.5, ct))
    cfgs.append(("z", "z", "z", CS, 14, False, 0.5, 3))
    cfgs.append(("c", "c", "c", CS, 6, True, 0.5, 3))
    csv = open(a.csv, "w") if a.csv else None
    if csv:
        csv.write("phi,m,n,k,function,relerr_max,relerr_med,TFLOPS,total_time [sec],conv_64f_2_8i,cublasGemmEx,conv_32i_2_8u,inverse_scaling,\n")

    def csv_rows(row):
        if not csv or "error" in row:
            return
        pre = f"{row['phi']:e},{row['m']},{row['n']},{row['k']},"
        tag = row["types"] + ("-" + row["computeType"] if row["computeType"] != "real" else "")
        ph = row["phases_ms"]
        csv.write(pre + f"{row['function']}[{tag}],{row['relerr_max']:e},{row['relerr_med']:e},{row['TFLOPS']:e},{row['total_time_ms'] / 1e3:e},"
                        f"{ph['scaling'] / 1e3:e},{ph['int8_gemm_residues'] / 1e3:e},{0.0:e},{ph['crt'] / 1e3:e},\n")
        ref = row.get("reference")
        if ref and "error" not in ref:
            rp = ref["phases_ms"]
            csv.write(pre + f"REF-{row['function']}[{tag}],{ref['relerr_max']:e},{ref['relerr_med']:e},{ref['TFLOPS']:e},{ref['total_time_ms'] / 1e3:e},"
                            f"{rp[0] / 1e3:e},{rp[1] / 1e3:e},{rp[2] / 1e3:e},{rp[3] / 1e3:e},\n")
        nat = row.get("native")
        if nat:
            name = {"d": "DGEMM", "s": "SGEMM", "z": "ZGEMM", "c": "CGEMM"}[row["types"][-1]]
            csv.write(pre + f"{name}[{tag}],{nat['relerr_max']:e},{nat['relerr_med']:e},{nat['TFLOPS']:e},,,,,,\n")
        csv.flush()

    with open(a.out, "w") as f:
        for c in cfgs:
            try:
                row = one(*c, reps=reps, with_ref=not a.no_ref)
            except Exception as e:
                row = {"config": list(map(str, c)), "error": str(e)[:300]}
            csv_rows(row)
            line = json.dumps(row)
            print(line, flush=True)
            f.write(line + "\n")
            f.flush()
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
