#!/usr/bin/env python3
"""Per-kernel summary of an `ncu --set full` raw CSV (ncu -i x.ncu-rep --page raw --csv) -> profiles/r01_ncu_traffic.json,
the file bench.py reads `roofline.traffic` from.  usage: tools/ncu_traffic.py RAW.csv OUT.json "source note" [keep.json]
(entries of keep.json whose kernel is absent from the new capture are carried over)."""
import csv
import json
import re
import sys

raw, out, note = sys.argv[1], sys.argv[2], sys.argv[3]
keep = json.load(open(sys.argv[4])) if len(sys.argv) > 4 else {}
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name, scale_to=None):
    if name not in col or r[col[name]] == "":
        return None
    v = float(r[col[name]].replace(",", ""))
    u = units[col[name]]
    if scale_to == "bytes":
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
    if scale_to == "ms":
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1, "second": 1e3}.get(u, 1)
    if scale_to == "GHz":
        v *= {"Hz": 1e-9, "Khz": 1e-6, "Mhz": 1e-3, "Ghz": 1, "hz": 1e-9, "cycle/nsecond": 1, "cycle/second": 1e-9}.get(u, 1)
    return v


res = dict(keep)
res.pop("_source", None)
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    m = re.search(r"(\w+_kernel)", name)
    if not m:
        continue
    rd, wr = val(r, "dram__bytes_read.sum", "bytes"), val(r, "dram__bytes_write.sum", "bytes")
    res[m.group(1)] = {
        "dram_bytes_per_launch": (rd or 0) + (wr or 0), "dram_read_GB": (rd or 0) / 1e9, "dram_write_GB": (wr or 0) / 1e9,
        "ncu_duration_ms": val(r, "gpu__time_duration.sum", "ms"),
        "tensor_imma_pipe_active_pct": val(r, "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warp_instructions": val(r, "smsp__inst_executed.sum"),
        "l2_hit_pct": val(r, "lts__t_sector_hit_rate.pct"),
        "sm_clock_GHz": val(r, "gpc__cycles_elapsed.avg.per_second", "GHz"),
        "registers_per_thread": val(r, "launch__registers_per_thread"),
    }
res["_source"] = note
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
