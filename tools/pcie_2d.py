"""PCIe copy rates: contiguous vs 2-D (row blocks of a column-major matrix) copies, H2D and D2H, pinned host memory."""
import ctypes as C, torch, time
rt = C.CDLL("libcudart.so.12")
rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
n = 16384
h = torch.empty((n, n), dtype=torch.float64, pin_memory=True); h.fill_(1.0)
d = torch.empty((n, n), dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
tot = n * n * 8
print("contiguous H2D  %.1f GB/s" % (tot / t(lambda: rt.cudaMemcpyAsync(d.data_ptr(), h.data_ptr(), tot, 1, st)) / 1e6))
print("contiguous D2H  %.1f GB/s" % (tot / t(lambda: rt.cudaMemcpyAsync(h.data_ptr(), d.data_ptr(), tot, 2, st)) / 1e6))
for rows in (256, 1024, 2048, 4096):
    w = rows * 8
    def h2d():
        for r0 in range(0, n, rows):
            rt.cudaMemcpy2DAsync(d.data_ptr() + r0 * 8, n * 8, h.data_ptr() + r0 * 8, n * 8, w, n, 1, st)
    def d2h():
        for r0 in range(0, n, rows):
            rt.cudaMemcpy2DAsync(h.data_ptr() + r0 * 8, n * 8, d.data_ptr() + r0 * 8, n * 8, w, n, 2, st)
    print("2-D blocks of %5d rows (%6d B pieces): H2D %.1f GB/s   D2H %.1f GB/s" % (rows, w, tot / t(h2d) / 1e6, tot / t(d2h) / 1e6))
# both directions at once on two streams
s2 = torch.cuda.Stream()
def both():
    rt.cudaMemcpyAsync(d.data_ptr(), h.data_ptr(), tot, 1, st)
    rt.cudaMemcpyAsync(h.data_ptr(), d.data_ptr(), tot // 2, 2, s2.cuda_stream)
ms = t(both)
print("H2D 2.1 GB + D2H 1.07 GB concurrently: %.1f ms (H2D alone would be %.1f ms)" % (ms, tot / 55e6))
