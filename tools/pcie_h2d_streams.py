"""H2D rate beside a saturating D2H as a function of the number of copy streams the H2D is spread over (and the reverse).
The end of the host-buffer call moves B column blocks in and C column strips out at the same time; one stream each way gave
43 GB/s in, 43 GB/s out there.  usage: pcie_h2d_streams.py  -> one JSON line per case"""
import ctypes as C, json, torch
rt = C.CDLL("libcudart.so.12")
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
GB = 1 << 30
hI = torch.empty(2 * GB, dtype=torch.uint8, pin_memory=True); hI.fill_(1)
hO = torch.empty(2 * GB, dtype=torch.uint8, pin_memory=True)
dI = torch.empty(2 * GB, dtype=torch.uint8, device="cuda"); dO = torch.ones(2 * GB, dtype=torch.uint8, device="cuda")
ins = [torch.cuda.Stream() for _ in range(4)]
outs = [torch.cuda.Stream() for _ in range(4)]
CHUNK = 64 << 20    # the call's copies are 32 - 100 MB each


def run(n_in, n_out, total=2 * GB):
    """`total` bytes each way in CHUNK-sized copies, round-robin over n_in / n_out streams (0: that direction idle)"""
    res = {}
    for attempt in range(2):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e0.record()
        for s in ins[:n_in] + outs[:n_out]:
            s.wait_event(e0)
        for i in range(total // CHUNK):
            off = i * CHUNK
            if n_in:
                rt.cudaMemcpyAsync(dI.data_ptr() + off, hI.data_ptr() + off, CHUNK, 1, ins[i % n_in].cuda_stream)
            if n_out:
                rt.cudaMemcpyAsync(hO.data_ptr() + off, dO.data_ptr() + off, CHUNK, 2, outs[i % n_out].cuda_stream)
        ei = [torch.cuda.Event(enable_timing=True) for _ in range(n_in)]
        eo = [torch.cuda.Event(enable_timing=True) for _ in range(n_out)]
        for e, s in zip(ei, ins): e.record(s)
        for e, s in zip(eo, outs): e.record(s)
        torch.cuda.synchronize()
        t_in = max([e0.elapsed_time(e) for e in ei], default=0.0)
        t_out = max([e0.elapsed_time(e) for e in eo], default=0.0)
        res = {"in_streams": n_in, "out_streams": n_out, "h2d_ms": round(t_in, 2), "d2h_ms": round(t_out, 2),
               "both_done_ms": round(max(t_in, t_out), 2),
               "h2d_GBps": round(total / t_in / 1e6, 1) if n_in else None, "d2h_GBps": round(total / t_out / 1e6, 1) if n_out else None}
    return res


for n_in, n_out in ((1, 0), (0, 1), (1, 1), (2, 1), (4, 1), (1, 2), (2, 2), (4, 4)):
    print(json.dumps(run(n_in, n_out)), flush=True)
