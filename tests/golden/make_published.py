#!/usr/bin/env python3
"""Extract the emulation rows of the reference's PUBLISHED accuracy table into a small fixture.

    python tests/golden/make_published.py        # build container only (reads /root/reference)

Source: GEMMul8/testing/results_in_paper/oz2_results_d_accuracy_NVIDIA_GH200_480GB_2025-04-09_02-40-54.csv
(test_double accuracy_check: m = n = 1024, k in {1024..16384}, phi in {0.5, 1, 2, 3, 4}, 2..20 moduli, fast and
accurate mode; the same numbers are in the A100 file: emulation results do not depend on the GPU).  Kept: the
OS2-fast / OS2-accu rows for k <= 4096 -- the known answers tests/test_parity_gpu.py reproduces digit for digit.
"""
import os

SRC = "/root/reference/GEMMul8/testing/results_in_paper/oz2_results_d_accuracy_NVIDIA_GH200_480GB_2025-04-09_02-40-54.csv"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "published_d_accuracy_GH200.csv")

keep = []
for i, line in enumerate(open(SRC)):
    f = line.strip().split(",")
    if i == 0:
        keep.append(line.strip())
    elif f[1].startswith("OS2-") and int(f[1].split("k=")[1].rstrip(")")) <= 4096:
        keep.append(line.strip())
open(DST, "w").write("\n".join(keep) + "\n")
print(len(keep) - 1, "rows ->", DST)
