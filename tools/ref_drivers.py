#!/usr/bin/env python3
"""Run the reference's own, UNMODIFIED benchmark drivers (GEMMul8/testing/test_*.cu, built by oracle/Makefile
into oracle/_ref/drivers/) against OUR library and against the reference library, and compare the CSVs they write
(`oz2_results_*_accuracy_*.csv` / `oz2_results_*_time_*.csv`, GEMMul8/testing/test_double.cu:65-213).

  run      (GPU box)  tools/ref_drivers.py run --out gpurun_out/refdrivers [--drivers test_double,...] [--checks accuracy_check,flops_check]
  compare  (anywhere) tools/ref_drivers.py compare --out gpurun_out/refdrivers [--paper DIR] [--md profiles/....md]

`compare` reports, per driver: how many emulation cells of the accuracy table are string-identical between the two
libraries (the drivers print 7 significant digits) and -- where the reference's results_in_paper/ directory is
readable (build container only) -- against the published GH200 table; and the TFLOPS of every OS2-* row of the
time table, ours against the reference library on the same GPU.
"""
import argparse
import glob
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRV = os.path.join(ROOT, "oracle", "_ref", "drivers")
ALL = ["test_double", "test_float", "test_mixed_double", "test_mixed_float", "test_float_complex"]


def run(a):
    rows = []
    for d in a.drivers.split(","):
        for lib in ("ours", "ref"):
            exe = os.path.join(DRV, f"{d}_{lib}")
            if not os.path.exists(exe):
                print(f"missing {exe} (make -C oracle drivers, in the build container)", file=sys.stderr)
                continue
            wd = os.path.join(a.out, d, lib)
            os.makedirs(wd, exist_ok=True)
            for f in glob.glob(os.path.join(wd, "oz2_results_*.csv")):
                os.unlink(f)
            for check in a.checks.split(","):
                t0 = time.time()
                with open(os.path.join(wd, check + ".stdout"), "w") as so:
                    r = subprocess.run([exe, check], cwd=wd, stdout=so, stderr=subprocess.STDOUT, timeout=a.timeout)
                rows.append({"driver": d, "lib": lib, "check": check, "rc": r.returncode, "seconds": round(time.time() - t0, 1)})
                print(json.dumps(rows[-1]), flush=True)
    json.dump(rows, open(os.path.join(a.out, "runs.json"), "w"), indent=1)
    return 0 if all(r["rc"] == 0 for r in rows) else 1


def read_csv(path):
    out = []
    for line in open(path):
        f = [x.strip() for x in line.rstrip("\n").split(",")]
        if len(f) > 2:
            out.append(f)
    return out


def one_csv(wd, kind):
    c = sorted(glob.glob(os.path.join(wd, f"oz2_results_*_{kind}_*.csv")))
    return read_csv(c[-1]) if c else None


def is_emulation(label):
    return label.startswith("OS2-")


def accuracy_table(rows):
    """{(phi, label): [cells]} of an accuracy CSV (header: phi,function,<moduli...>)."""
    t = {}
    for f in rows[1:]:
        t[(float(f[0]), f[1])] = [c for c in f[2:] if c != ""]
    return t, [c for c in rows[0][2:] if c != ""]


def time_table(rows):
    """{(m, label): row dict} of a time CSV."""
    hdr = rows[0]
    t = {}
    for f in rows[1:]:
        d = dict(zip(hdr, f))
        t[(int(d["m"]), d["function"])] = d
    return t


PAPER = {"test_double": "oz2_results_d_accuracy_NVIDIA_GH200_480GB_2025-04-09_02-40-54.csv",
         "test_float": "oz2_results_f_accuracy_NVIDIA_GH200_480GB_2025-04-09_01-42-47.csv"}


def compare(a):
    md = ["# The reference's own drivers (unmodified GEMMul8/testing/test_*.cu) linked against libgemmul8_b200.so and against the reference library, same B200",
          ""]
    summary = {}
    for d in a.drivers.split(","):
        ours_wd, ref_wd = os.path.join(a.out, d, "ours"), os.path.join(a.out, d, "ref")
        oa, ra = one_csv(ours_wd, "accuracy"), one_csv(ref_wd, "accuracy")
        s = summary.setdefault(d, {})
        if oa and ra:
            to, moduli = accuracy_table(oa)
            tr, _ = accuracy_table(ra)
            same = tot = 0
            diffs = []
            for key, cells in to.items():
                if not is_emulation(key[1]) or key not in tr:
                    continue
                for nm, x, y in zip(moduli, cells, tr[key]):
                    tot += 1
                    same += x == y
                    if x != y:
                        diffs.append((key, nm, x, y))
            s["accuracy_cells"] = tot
            s["accuracy_cells_identical_to_reference_library"] = same
            md += [f"## {d}: accuracy_check", "",
                   f"* emulation cells (phi x k x num_moduli x fast/accurate): **{same} of {tot} string-identical** to the reference library's output on the same GPU"]
            for key, nm, x, y in diffs[:12]:
                md.append(f"  * differs: phi={key[0]} {key[1]} N={nm}: ours {x}, reference {y}")
            paper = os.path.join(a.paper, PAPER.get(d, "-"))
            if os.path.exists(paper):
                tp, pm = accuracy_table(read_csv(paper))
                ps = pt = 0
                pdiff = []
                for key, cells in to.items():
                    if not is_emulation(key[1]) or key not in tp:
                        continue
                    prow = dict(zip(pm, tp[key]))
                    for nm, x in zip(moduli, cells):
                        if nm in prow:
                            pt += 1
                            ps += x == prow[nm]
                            if x != prow[nm]:
                                pdiff.append((key, nm, x, prow[nm]))
                s["accuracy_cells_in_published_table"] = pt
                s["accuracy_cells_identical_to_published_GH200_table"] = ps
                md.append(f"* against the PUBLISHED table ({PAPER[d]}, GH200, 2025-04-09): **{ps} of {pt} cells string-identical**")
                for key, nm, x, y in pdiff[:12]:
                    md.append(f"  * differs: phi={key[0]} {key[1]} N={nm}: ours {x}, published {y}")
            md.append("")
        ot, rt = one_csv(ours_wd, "time"), one_csv(ref_wd, "time")
        if ot and rt:
            to, tr = time_table(ot), time_table(rt)
            md += [f"## {d}: flops_check (the driver's own timing: average of 100 synchronised calls, phi = 0.5)", "",
                   "| m=n=k | function | relerr_max ours | relerr_max reference | TFLOPS ours | TFLOPS reference | ratio |", "|---|---|---|---|---|---|---|"]
            ratios = []
            err_same = err_tot = 0
            for key in to:
                o, r = to[key], tr.get(key)
                if r is None:
                    continue
                try:
                    fo, fr = float(o["TFLOPS"]), float(r["TFLOPS"])
                except ValueError:
                    continue
                if is_emulation(key[1]):
                    ratios.append(fo / fr)
                    err_tot += 1
                    err_same += o["relerr_max"] == r["relerr_max"] and o["relerr_med"] == r["relerr_med"]
                show = (not is_emulation(key[1])) or key[1].rsplit("-", 1)[-1] in a.show.split(",")
                if show:
                    md.append(f"| {key[0]} | {key[1]} | {o['relerr_max']} | {r['relerr_max']} | {fo:.1f} | {fr:.1f} | {fo / fr:.2f} |")
            if ratios:
                ratios.sort()
                s["time_rows"] = len(ratios)
                s["tflops_ratio_min_median_max"] = [round(ratios[0], 3), round(ratios[len(ratios) // 2], 3), round(ratios[-1], 3)]
                s["time_rows_with_identical_errors"] = err_same
                md += ["", f"* {len(ratios)} emulation rows (sizes x num_moduli x mode): ours / reference TFLOPS min {ratios[0]:.2f}, median "
                       f"{ratios[len(ratios) // 2]:.2f}, max {ratios[-1]:.2f}; relerr_max and relerr_med string-identical in {err_same} of {err_tot} rows", ""]
    print(json.dumps(summary, indent=1))
    if a.md:
        md += ["## summary", "", "```json", json.dumps(summary, indent=1), "```", ""]
        open(a.md, "w").write("\n".join(md))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cmd", choices=["run", "compare"])
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "refdrivers"))
    ap.add_argument("--drivers", default=",".join(ALL))
    ap.add_argument("--checks", default="accuracy_check,flops_check")
    ap.add_argument("--timeout", type=int, default=1500)
    ap.add_argument("--paper", default="/root/reference/GEMMul8/testing/results_in_paper")
    ap.add_argument("--md", default=None)
    ap.add_argument("--show", default="6,8,10,12,14,15,16,18,20", help="num_moduli values listed row by row in the markdown table")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    return run(a) if a.cmd == "run" else compare(a)


if __name__ == "__main__":
    sys.exit(main())
