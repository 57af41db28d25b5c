#!/usr/bin/env python3
"""Phase split and back-to-back call time at small and medium sizes (where launch latency and grid sizing, not the
tensor pipe, decide): tools/small_sizes.py [--sizes 1024,2048,4096,8192] [--moduli 14]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gemmul8_b200 as g


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1024,2048,4096,8192")
    ap.add_argument("--moduli", type=int, default=14)
    ap.add_argument("--accurate", action="store_true")
    ap.add_argument("--flags", type=int, default=0, help="extra GEMMUL8_FLAG_* bits (64 = the single-kernel GEMM + CRT)")
    ap.add_argument("--shapes", default="", help="m x n x k triples instead of squares, e.g. 16384x16384x1024,16384x16384x512 (HPL-like updates)")
    a = ap.parse_args()
    N, fast = a.moduli, not a.accurate
    shapes = [tuple(int(v) for v in t.split("x")) for t in a.shapes.split(",") if t] or [(int(x),) * 3 for x in a.sizes.split(",")]
    for (m, n, k) in shapes:
        S = max(m, n, k)
        A = g.phi_matrix(m, k, 0.5, torch.float64)
        B = g.phi_matrix(k, n, 0.5, torch.float64, seed=7)
        C = torch.zeros((n, m), dtype=torch.float64, device="cuda")
        work = torch.empty(g.workSize(m, n, k, N), dtype=torch.uint8, device="cuda")
        args = g.make_args(0, 0, m, n, k, 1.0, A, m, B, k, 0.0, C, m, N, fast, work, flags=a.flags)
        import ctypes
        call = lambda: g.lib().gemmul8_b200_gemm(ctypes.byref(args))
        for _ in range(5):
            call()
        torch.cuda.synchronize()
        reps = 200 if S <= 2048 else 50
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            call()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        ph = [0.0] * 4
        for _ in range(10):
            t = g.gemm(None, 0, 0, m, n, k, 1.0, A, m, B, k, 0.0, C, m, N, fast, work, flags=g.FLAG_TIMERS | a.flags)
            ph = [x + y / 10 / 1e3 for x, y in zip(ph, t)]
        print(json.dumps({"m": m, "n": n, "k": k, "moduli": N, "fast": fast, "us_per_call_back_to_back": round(us, 1),
                          "TFLOPS": round(2.0 * m * n * k / us / 1e6, 1),
                          "phases_us": {"scaling": round(ph[0], 1), "gemm": round(ph[1], 1), "crt": round(ph[3], 1)},
                          "gemm_ideal_us_at_2842_TOPS": round(2.0 * N * m * n * k / 2842e12 * 1e6, 1)}), flush=True)


if __name__ == "__main__":
    main()
