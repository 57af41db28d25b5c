"""A/B timing of library builds inside one GPU visit: each variant runs in its own process (GEMMUL8_B200_LIB), interleaved
rounds, same inputs.  usage: ab_time.py name=path[,name=path...] "m,n,k,N;..." [rounds]  -> one JSON line per (shape, variant)"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    import torch
    import gemmul8_b200 as g
    g.init()
    out = {}
    for spec in sys.argv[2].split(";"):
        m, n, k, N = (int(x) for x in spec.split(",")[:4])
        flags = int(spec.split(",")[4]) if len(spec.split(",")) > 4 else 0
        A = g.phi_matrix(m, k, 0.5, torch.float64)
        B = g.phi_matrix(k, n, 0.5, torch.float64, seed=7)
        C = torch.zeros((n, m), dtype=torch.float64, device="cuda")
        work = torch.empty(g.workSize(m, n, k, N), dtype=torch.uint8, device="cuda")
        for _ in range(3):
            g.gemm(None, 0, 0, m, n, k, 1.0, A, m, B, k, 0.0, C, m, N, True, work, flags=flags)
        torch.cuda.synchronize()
        reps = max(5, min(40, int(2e12 / (2.0 * m * n * k * N / 14 + 1))))
        g.phase_log_collect()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.gemm(None, 0, 0, m, n, k, 1.0, A, m, B, k, 0.0, C, m, N, True, work, flags=flags | g.FLAG_PHASE_LOG)
        e1.record()
        torch.cuda.synchronize()
        ph, _ = g.phase_log_collect()
        out[spec] = {"us": e0.elapsed_time(e1) / reps * 1e3, "phases_us": [x / reps / 1e3 for x in ph], "sum": float(C.double().sum())}
        del A, B, C, work
    print(json.dumps(out))
    sys.exit(0)

variants = [v.split("=", 1) for v in sys.argv[1].split(",")]
shapes = sys.argv[2]
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 2
acc = {}
for r in range(rounds):
    for name, path in variants:
        env = dict(os.environ)
        path, *extra = path.split(":")           # name=path[:ENV=VAL]...
        for kv in extra:
            env[kv.split("=")[0]] = kv.split("=")[1]
        if path != "default":
            env["GEMMUL8_B200_LIB"] = os.path.join(ROOT, path)
        p = subprocess.run([sys.executable, __file__, "--child", shapes], env=env, capture_output=True, text=True, timeout=900)
        if p.returncode != 0:
            print(json.dumps({"variant": name, "error": p.stderr[-400:]}), flush=True)
            continue
        res = json.loads(p.stdout.strip().splitlines()[-1])
        for spec, d in res.items():
            acc.setdefault(spec, {}).setdefault(name, []).append(d)
for spec, byv in acc.items():
    for name, runs in byv.items():
        best = min(runs, key=lambda d: d["us"])
        print(json.dumps({"shape": spec, "variant": name, "us_best": round(best["us"], 1), "us_all": [round(d["us"], 1) for d in runs],
                          "phases_us": [round(x, 1) for x in best["phases_us"]], "sum": best["sum"]}), flush=True)
