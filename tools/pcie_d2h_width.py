"""D2H rate of strided (2-D) copies of C row strips as a function of the run length, alone and beside a saturating H2D, on one
or two copy streams.  The host-buffer call returns row strips of C as 2-D copies whose runs are (block rows) x 8 bytes.
usage: pcie_d2h_width.py  -> one JSON line per case"""
import ctypes as C, json, torch
rt = C.CDLL("libcudart.so.12")
rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
n = 16384
hA = torch.empty((n, n), dtype=torch.float64, pin_memory=True); hA.fill_(1.0)
hC = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
dA = torch.empty((n, n), dtype=torch.float64, device="cuda"); dC = torch.zeros_like(dA)
s_in = torch.cuda.Stream()
outs = [torch.cuda.Stream() for _ in range(4)]

def d2h(rows, streams, cols=12288):
    """rows [0, rows) of `cols` columns, split by columns over `streams` copy streams"""
    per = cols // streams
    for i in range(streams):
        off = i * per * n * 8
        rt.cudaMemcpy2DAsync(hC.data_ptr() + off, n * 8, dC.data_ptr() + off, n * 8, rows * 8, per, 2, outs[i].cuda_stream)

def run(rows, streams, busy):
    cols = 12288
    reps = max(1, 4096 // rows // 2)
    for attempt in range(2):
        torch.cuda.synchronize()
        if busy:
            for _ in range(3):
                rt.cudaMemcpyAsync(dA.data_ptr(), hA.data_ptr(), n * n * 8, 1, s_in.cuda_stream)
        e0 = torch.cuda.Event(enable_timing=True)
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(streams)]
        e0.record(outs[0])
        for i in range(1, streams):
            outs[i].wait_event(e0)
        for _ in range(reps):
            d2h(rows, streams, cols)
        for i in range(streams):
            ends[i].record(outs[i])
        torch.cuda.synchronize()
        ms = max(e0.elapsed_time(e) for e in ends)
    return rows * 8 * cols * reps / ms / 1e6

for rows in (256, 512, 1024, 2048, 4096, 16384):
    for streams in (1, 2, 4):
        print(json.dumps({"run_bytes": rows * 8, "streams": streams, "GBps_alone": round(run(rows, streams, False), 1),
                          "GBps_beside_h2d": round(run(rows, streams, True), 1)}), flush=True)
