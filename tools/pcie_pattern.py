"""The transfer pattern of gemmul8_b200_gemm_host at 16384^3 without any compute: H2D of A row blocks (2-D copies) and B column
blocks (contiguous) in the 12-block shrinking order, alone and with the D2H of the C strips running beside it."""
import ctypes as C, json, torch
rt = C.CDLL("libcudart.so.12")
rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
n = 16384
hA = torch.empty((n, n), dtype=torch.float64, pin_memory=True); hA.fill_(1.0)
hB = torch.empty((n, n), dtype=torch.float64, pin_memory=True); hB.fill_(1.0)
hC = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
dA = torch.empty((n, n), dtype=torch.float64, device="cuda"); dB = torch.empty_like(dA); dC = torch.zeros_like(dA)
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
w = [3, 3, 3, 3, 2, 2, 2, 2, 1, 1, 1, 1]
tiles = n // 256
b = [0]
acc = 0
for x in w:
    acc += x
    b.append(tiles * acc // 24 * 256)
b[-1] = n

def h2d(uniform=False):
    bb = [n * i // 8 for i in range(9)] if uniform else b
    for i in range(len(bb) - 1):
        r0, r1 = bb[i], bb[i + 1]
        rt.cudaMemcpy2DAsync(dA.data_ptr() + r0 * 8, n * 8, hA.data_ptr() + r0 * 8, n * 8, (r1 - r0) * 8, n, 1, s_in.cuda_stream)   # rows r0..r1 of A
        rt.cudaMemcpyAsync(dB.data_ptr() + r0 * n * 8, hB.data_ptr() + r0 * n * 8, (r1 - r0) * n * 8, 1, s_in.cuda_stream)            # columns of B

def d2h():
    for i in range(len(b) - 1):
        r0, r1 = b[i], b[i + 1]
        # column strip: rows [0, r1) of columns [r0, r1); row strip: rows [r0, r1) of columns [0, r0)
        rt.cudaMemcpy2DAsync(hC.data_ptr() + r0 * n * 8, n * 8, dC.data_ptr() + r0 * n * 8, n * 8, r1 * 8, r1 - r0, 2, s_out.cuda_stream)
        if r0:
            rt.cudaMemcpy2DAsync(hC.data_ptr() + r0 * 8, n * 8, dC.data_ptr() + r0 * 8, n * 8, (r1 - r0) * 8, r0, 2, s_out.cuda_stream)

def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def contiguous():
    rt.cudaMemcpyAsync(dA.data_ptr(), hA.data_ptr(), n * n * 8, 1, s_in.cuda_stream)
    rt.cudaMemcpyAsync(dB.data_ptr(), hB.data_ptr(), n * n * 8, 1, s_in.cuda_stream)

out = {"h2d_contiguous_ms": t(contiguous), "h2d_8_equal_blocks_ms": t(lambda: h2d(True)), "h2d_12_shrinking_blocks_ms": t(h2d), "d2h_strips_ms": t(d2h),
       "h2d_and_d2h_together_ms": t(lambda: (h2d(), d2h()))}
print(json.dumps(out))
