"""gemmul8_b200 -- host-side mirror of the reference's operator interface for the hot path.

The product is the C-ABI shared library ``libgemmul8_b200.so`` (include/gemmul8_b200.h); this
module is a thin ctypes binding with the reference's names and argument order
(``gemmul8::workSize`` / ``gemmul8::gemm``, GEMMul8/include/gemmul8.hpp:18-47) so that tests read
like the reference's own drivers.  PyTorch is used for device memory and streams only.

There is no CPU fallback: importing works without a GPU (so the ABI can be inspected), but
``gemm`` raises ``Gemmul8Error`` if the CUDA path cannot run, and loading fails loudly if the
extension has not been built (``python -m`` ``__graft_entry__`` / ``build.py``).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# (GEMMUL8_B200_LIB: another build of the same library, for A/B timing of kernel variants inside one GPU visit)
LIB_PATH = os.environ.get("GEMMUL8_B200_LIB") or os.path.join(_HERE, "libgemmul8_b200.so")
AUX_PATH = os.path.join(_HERE, "libgemmul8_b200_aux.so")

# gemmul8::computeType_t (GEMMul8/include/gemmul8.hpp:7-12)
REAL_DEFAULT, COMPLEX_BIG_MATRIX_ENCODE, COMPLEX_CLASSIC_MULT, COMPLEX_KARATSUBA_MULT = 0, 1, 2, 3
# cublasOperation_t
OP_N, OP_T, OP_C = 0, 1, 2
F32, F64, C32, C64 = 0, 1, 2, 3
FLAG_TIMERS, FLAG_STAGE_SCALING, FLAG_STAGE_RESIDUES, FLAG_FUSED_CRT, FLAG_GEMM_SIMT, FLAG_HOST_SERIAL, FLAG_STRIPS = 1, 1 << 4, 1 << 5, 1 << 6, 1 << 8, 1 << 9, 1 << 10
FLAG_ONLY_SCALE_A, FLAG_SKIP_SCALE_A, FLAG_PHASE_LOG = 1 << 11, 1 << 12, 1 << 13
FLAG_ONLY_BOUND, FLAG_SKIP_BOUND = 1 << 14, 1 << 15
FLAG_DEVICE_SCALARS, FLAG_EXCLUSIVE_SMS = 1 << 16, 1 << 17

EXPORTED_SYMBOLS = (
    "gemmul8_b200_worksize", "gemmul8_b200_work_layout", "gemmul8_b200_gemm", "gemmul8_b200_host_scratch_size",
    "gemmul8_b200_gemm_host", "gemmul8_b200_gemm_part",
    "gemmul8_b200_worksize_blocked", "gemmul8_b200_worksize_blocked_complex", "gemmul8_b200_plan_blocks", "gemmul8_b200_gemm_blocked", "gemmul8_b200_product_i32", "gemmul8_b200_modulus", "gemmul8_b200_crt_weight",
    "gemmul8_b200_phase_log_collect", "gemmul8_b200_launch_count", "gemmul8_b200_last_error", "gemmul8_b200_version",
    "gemmul8_b200_init", "gemmul8_b200_set_option", "gemmul8_b200_get_option",
)


class Gemmul8Error(RuntimeError):
    pass


class Args(C.Structure):
    """gemmul8_b200_args (include/gemmul8_b200.h)."""
    _fields_ = [
        ("op_A", C.c_int), ("op_B", C.c_int),
        ("m", C.c_size_t), ("n", C.c_size_t), ("k", C.c_size_t),
        ("alpha", C.c_void_p),
        ("A", C.c_void_p), ("lda", C.c_size_t),
        ("B", C.c_void_p), ("ldb", C.c_size_t),
        ("beta", C.c_void_p),
        ("C", C.c_void_p), ("ldc", C.c_size_t),
        ("num_moduli", C.c_uint), ("fastmode", C.c_int),
        ("work", C.c_void_p), ("compute_type", C.c_int),
        ("dtype_A", C.c_int), ("dtype_B", C.c_int), ("dtype_C", C.c_int),
        ("stream", C.c_void_p), ("flags", C.c_uint),
        ("timers_ns", C.c_double * 4),
    ]


class Layout(C.Structure):
    """gemmul8_b200_layout: the reference's carve of `work` (GEMMul8/src/gemmul8.cu:229-234)."""
    _fields_ = [(n, C.c_size_t) for n in (
        "lda8i", "m_pad", "sizeA", "sizeB", "sizeC", "off_A8i", "off_A8i_imag", "off_B8i", "off_B8i_imag",
        "off_C8u", "off_C8u_imag", "off_C32i", "off_C32i_imag", "off_sftA", "off_sftB", "total")]


_lib = None
_aux = None


def lib():
    """The product library; raises if it has not been built (no silent fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Gemmul8Error(f"{LIB_PATH} is missing: run `python __graft_entry__.py` (build) first; "
                               "there is no CPU or PyTorch fallback for this path")
        L = C.CDLL(LIB_PATH)
        L.gemmul8_b200_worksize.restype = C.c_size_t
        L.gemmul8_b200_worksize.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint, C.c_int]
        L.gemmul8_b200_work_layout.restype = C.c_int
        L.gemmul8_b200_work_layout.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint, C.c_int, C.POINTER(Layout)]
        L.gemmul8_b200_gemm.restype = C.c_int
        L.gemmul8_b200_gemm.argtypes = [C.POINTER(Args)]
        L.gemmul8_b200_host_scratch_size.restype = C.c_size_t
        L.gemmul8_b200_host_scratch_size.argtypes = [C.POINTER(Args)]
        L.gemmul8_b200_gemm_host.restype = C.c_int
        L.gemmul8_b200_gemm_host.argtypes = [C.POINTER(Args), C.c_void_p]
        L.gemmul8_b200_gemm_part.restype = C.c_int
        L.gemmul8_b200_gemm_part.argtypes = [C.POINTER(Args), C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t]
        L.gemmul8_b200_worksize_blocked.restype = C.c_size_t
        L.gemmul8_b200_worksize_blocked.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint, C.c_size_t, C.c_size_t]
        L.gemmul8_b200_worksize_blocked_complex.restype = C.c_size_t
        L.gemmul8_b200_worksize_blocked_complex.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint, C.c_int, C.c_size_t, C.c_size_t]
        L.gemmul8_b200_plan_blocks.restype = C.c_int
        L.gemmul8_b200_plan_blocks.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint, C.c_size_t] + [C.POINTER(C.c_size_t)] * 3
        L.gemmul8_b200_gemm_blocked.restype = C.c_int
        L.gemmul8_b200_gemm_blocked.argtypes = [C.POINTER(Args), C.c_size_t, C.c_size_t]
        L.gemmul8_b200_product_i32.restype = C.c_int
        L.gemmul8_b200_product_i32.argtypes = [C.POINTER(Args), C.c_uint, C.c_void_p, C.c_int]
        L.gemmul8_b200_modulus.restype = C.c_int
        L.gemmul8_b200_modulus.argtypes = [C.c_uint]
        L.gemmul8_b200_crt_weight.restype = C.c_double
        L.gemmul8_b200_crt_weight.argtypes = [C.c_uint, C.c_uint, C.c_int]
        L.gemmul8_b200_phase_log_collect.restype = C.c_int
        L.gemmul8_b200_phase_log_collect.argtypes = [C.POINTER(C.c_double * 4), C.POINTER(C.c_uint)]
        L.gemmul8_b200_launch_count.restype = C.c_ulonglong
        L.gemmul8_b200_init.restype = C.c_int
        L.gemmul8_b200_init.argtypes = [C.c_int]
        L.gemmul8_b200_set_option.restype = C.c_int
        L.gemmul8_b200_set_option.argtypes = [C.c_char_p, C.c_int]
        L.gemmul8_b200_get_option.restype = C.c_int
        L.gemmul8_b200_get_option.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
        L.gemmul8_b200_last_error.restype = C.c_char_p
        L.gemmul8_b200_version.restype = C.c_char_p
        _lib = L
    return _lib


def aux():
    """Measurement helpers (phi matrices, double-double truth); tests / bench only."""
    global _aux
    if _aux is None:
        if not os.path.exists(AUX_PATH):
            raise Gemmul8Error(f"{AUX_PATH} is missing: run the build first")
        L = C.CDLL(AUX_PATH)
        L.gemmul8_aux_phi_matrix.restype = C.c_int
        L.gemmul8_aux_phi_matrix.argtypes = [C.c_int, C.c_size_t, C.c_void_p, C.c_double, C.c_ulonglong, C.c_void_p]
        L.gemmul8_aux_dd_gemm.restype = C.c_int
        L.gemmul8_aux_dd_gemm.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p,
                                          C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        _aux = L
    return _aux


def _check(rc):
    if rc != 0:
        raise Gemmul8Error(f"gemmul8_b200 status {rc}: {lib().gemmul8_b200_last_error().decode()}")


def init(device=-1):
    """gemmul8_b200_init: read the tuning defaults and probe the SM placement table of `device` (default: current)."""
    _check(lib().gemmul8_b200_init(device))


def set_option(name, value):
    """Process-wide tuning / debug option (see include/gemmul8_b200.h)."""
    _check(lib().gemmul8_b200_set_option(name.encode(), int(value)))


def get_option(name):
    v = C.c_int()
    _check(lib().gemmul8_b200_get_option(name.encode(), C.byref(v)))
    return v.value


def workSize(m, n, k, num_moduli, computeType=REAL_DEFAULT):
    """gemmul8::workSize (GEMMul8/src/gemmul8.cu:129-147): bytes of device scratch; 0 for a bad computeType."""
    return lib().gemmul8_b200_worksize(m, n, k, num_moduli, computeType)


def work_layout(m, n, k, num_moduli, computeType=REAL_DEFAULT):
    out = Layout()
    _check(lib().gemmul8_b200_work_layout(m, n, k, num_moduli, computeType, C.byref(out)))
    return out


def _dtype_tag(t):
    import torch
    return {torch.float32: F32, torch.float64: F64, torch.complex64: C32, torch.complex128: C64}[t.dtype]


def _scalar_buf(value, tag):
    if tag == F32:
        return (C.c_float * 1)(float(value))
    if tag == F64:
        return (C.c_double * 1)(float(value))
    v = complex(value)
    return ((C.c_float if tag == C32 else C.c_double) * 2)(v.real, v.imag)


def make_args(op_A, op_B, m, n, k, alpha, A, lda, B, ldb, beta, Cmat, ldc, num_moduli, fastmode, work,
              computeType=REAL_DEFAULT, stream=None, flags=0):
    """Fill a gemmul8_b200_args from torch tensors (any shape; only data_ptr and dtype are used)."""
    import torch
    a = Args()
    a.op_A, a.op_B, a.m, a.n, a.k = op_A, op_B, m, n, k
    a.dtype_A, a.dtype_B, a.dtype_C = _dtype_tag(A), _dtype_tag(B), _dtype_tag(Cmat)
    if flags & FLAG_DEVICE_SCALARS:      # alpha, beta: one-element CUDA tensors of C's dtype, read on the device
        a._alpha, a._beta = alpha, beta
        a.alpha, a.beta = alpha.data_ptr(), beta.data_ptr()
    else:
        a._alpha = _scalar_buf(alpha, a.dtype_C)
        a._beta = _scalar_buf(beta, a.dtype_C)
        a.alpha, a.beta = C.addressof(a._alpha), C.addressof(a._beta)
    a.A, a.lda, a.B, a.ldb, a.C, a.ldc = A.data_ptr(), lda, B.data_ptr(), ldb, Cmat.data_ptr(), ldc
    a.num_moduli, a.fastmode, a.work, a.compute_type = num_moduli, int(bool(fastmode)), work.data_ptr(), computeType
    if stream is None and Cmat.is_cuda:
        stream = torch.cuda.current_stream(Cmat.device).cuda_stream
    a.stream, a.flags = stream or 0, flags
    a._keep = (A, B, Cmat, work)
    return a


def gemm(handle, op_A, op_B, m, n, k, alpha, A, lda, B, ldb, beta, Cmat, ldc, num_moduli, fastmode, work,
         computeType=REAL_DEFAULT, flags=0):
    """gemmul8::gemm<TA,TB,TC> (GEMMul8/include/gemmul8.hpp:29-47), same argument order.

    `handle` stands for the cuBLAS handle of the reference; only its stream matters here: pass None
    (torch's current stream), a torch.cuda.Stream, or a raw cudaStream_t integer.  A, B, C and work are
    CUDA tensors holding column-major data.  Returns the 4 phase times in ns (zeros unless FLAG_TIMERS).
    An unsupported computeType prints the reference's message and returns zeros with C untouched.
    """
    stream = None
    if handle is not None:
        stream = handle.cuda_stream if hasattr(handle, "cuda_stream") else int(handle)
    a = make_args(op_A, op_B, m, n, k, alpha, A, lda, B, ldb, beta, Cmat, ldc, num_moduli, fastmode, work,
                  computeType, stream, flags)
    rc = lib().gemmul8_b200_gemm(C.byref(a))
    if rc == 1:  # reference behaviour: message on stderr + zero timers, no exception
        return [0.0, 0.0, 0.0, 0.0]
    _check(rc)
    return list(a.timers_ns)


PART_SCALE_A, PART_SCALE_B, PART_PRODUCT = 1, 2, 4


def gemm_part(args, parts, row0, row1, col0, col1):
    """One or more steps (PART_*) of the real fast-mode path on rows [row0, row1) x columns [col0, col1) of the full
    problem described by `args` (from make_args); see gemmul8_b200_gemm_part in include/gemmul8_b200.h."""
    _check(lib().gemmul8_b200_gemm_part(C.byref(args), parts, row0, row1, col0, col1))


def workSizeBlocked(m, n, k, num_moduli, block_rows, block_cols):
    """Bytes of `work` for gemm_blocked with these block sizes (0 if they are not multiples of 256 / whole dimensions)."""
    return lib().gemmul8_b200_worksize_blocked(m, n, k, num_moduli, block_rows, block_cols)


def plan_blocks(m, n, k, num_moduli, max_bytes):
    """(block_rows, block_cols, work_bytes): the blocks with the least re-encoding whose workspace fits max_bytes."""
    mb, nb, wb = C.c_size_t(), C.c_size_t(), C.c_size_t()
    _check(lib().gemmul8_b200_plan_blocks(m, n, k, num_moduli, max_bytes, C.byref(mb), C.byref(nb), C.byref(wb)))
    return mb.value, nb.value, wb.value


def workSizeBlockedComplex(m, n, k, num_moduli, computeType, block_rows, block_cols):
    """Bytes of `work` for gemm_blocked on complex types (fast mode): workSize of one block."""
    return lib().gemmul8_b200_worksize_blocked_complex(m, n, k, num_moduli, computeType, block_rows, block_cols)


def gemm_blocked(handle, op_A, op_B, m, n, k, alpha, A, lda, B, ldb, beta, Cmat, ldc, num_moduli, fastmode, work,
                 block_rows, block_cols, flags=0, computeType=REAL_DEFAULT):
    """gemm() with a small workspace: C in blocks of block_rows x block_cols, `work` of workSizeBlocked() bytes
    (gemmul8_b200_gemm_blocked; real types).  Same result bits as gemm()."""
    stream = None
    if handle is not None:
        stream = handle.cuda_stream if hasattr(handle, "cuda_stream") else int(handle)
    a = make_args(op_A, op_B, m, n, k, alpha, A, lda, B, ldb, beta, Cmat, ldc, num_moduli, fastmode, work, computeType, stream, flags)
    _check(lib().gemmul8_b200_gemm_blocked(C.byref(a), block_rows, block_cols))
    return list(a.timers_ns)


def gemm_host(op_A, op_B, m, n, k, alpha, A, lda, B, ldb, beta, Cmat, ldc, num_moduli, fastmode, dev_scratch,
              computeType=REAL_DEFAULT, stream=None, flags=0):
    """Same call with (pinned) HOST tensors; copies in, computes on the GPU, copies C back, synchronises."""
    a = make_args(op_A, op_B, m, n, k, alpha, A, lda, B, ldb, beta, Cmat, ldc, num_moduli, fastmode, dev_scratch,
                  computeType, stream, flags)
    _check(lib().gemmul8_b200_gemm_host(C.byref(a), dev_scratch.data_ptr()))
    return list(a.timers_ns)


def host_scratch_size(op_A, op_B, m, n, k, A, lda, B, ldb, Cmat, ldc, num_moduli, computeType=REAL_DEFAULT):
    import torch
    a = make_args(op_A, op_B, m, n, k, 1.0, A, lda, B, ldb, 0.0, Cmat, ldc, num_moduli, True, torch.empty(0), computeType, 0)
    return lib().gemmul8_b200_host_scratch_size(C.byref(a))


def product_i32(args, j, out, imag_part=0):
    """Raw int32 product of modulus slice j (what cublasGemmEx writes at GEMMul8/src/gemmul8.cu:265)."""
    _check(lib().gemmul8_b200_product_i32(C.byref(args), j, out.data_ptr(), imag_part))


def work_views(work, L, num_moduli, m, n):
    """Typed torch views of the sub-buffers of `work` (uint8 tensor), as the reference lays them out."""
    import torch
    N = num_moduli
    v = {}
    v["A8i"] = work[L.off_A8i:L.off_A8i + N * L.sizeA].view(torch.int8).view(N, L.m_pad, L.lda8i)
    v["B8i"] = work[L.off_B8i:L.off_B8i + N * L.sizeB].view(torch.int8).view(N, n, L.lda8i)
    v["C8u"] = work[L.off_C8u:L.off_C8u + N * L.sizeC].view(N, L.sizeC)[:, :L.m_pad * n].view(N, n, L.m_pad)
    v["sftA"] = work[L.off_sftA:L.off_sftA + 2 * m].view(torch.int16)
    v["sftB"] = work[L.off_sftB:L.off_sftB + 2 * n].view(torch.int16)
    return v


def work_views_complex(work, L, num_moduli, m, n, k, computeType):
    """Typed views of `work` for complex calls, as the reference carves it
    (big matrix: GEMMul8/src/gemmul8.cu:659-664; CLASSIC / KARATSUBA: :806-815).
    big matrix : A8i (N, m2_pad, lda8i) with rows [Pr | -Pi] then [Pi | Pr]; B8i (N, n, lda8i) = [Qr | Qi];
                 C8u (N, n, m2_pad) with Re in rows < m and Im in rows m .. 2m-1
    otherwise  : A8i_real / A8i_imag, B8i_real / B8i_imag, C8u_real / C8u_imag"""
    import torch
    N = num_moduli
    v = {}

    def stack8(off, rows):
        return work[off:off + N * rows * L.lda8i].view(torch.int8).view(N, rows, L.lda8i)

    def stackC(off):
        return work[off:off + N * L.sizeC].view(N, L.sizeC)[:, :L.m_pad * n].view(N, n, L.m_pad)

    if computeType == COMPLEX_BIG_MATRIX_ENCODE:
        v["A8i"], v["B8i"], v["C8u"] = stack8(L.off_A8i, L.m_pad), stack8(L.off_B8i, n), stackC(L.off_C8u)
        v["C8u_real"], v["C8u_imag"] = v["C8u"][:, :, :m], v["C8u"][:, :, m:2 * m]
    else:
        v["A8i_real"], v["A8i_imag"] = stack8(L.off_A8i, L.m_pad), stack8(L.off_A8i_imag, L.m_pad)
        v["B8i_real"], v["B8i_imag"] = stack8(L.off_B8i, n), stack8(L.off_B8i_imag, n)
        v["C8u_real"], v["C8u_imag"] = stackC(L.off_C8u)[:, :, :m], stackC(L.off_C8u_imag)[:, :, :m]
    v["sftA"] = work[L.off_sftA:L.off_sftA + 2 * m].view(torch.int16)
    v["sftB"] = work[L.off_sftB:L.off_sftB + 2 * n].view(torch.int16)
    return v


def phase_log_collect():
    """(phase times in ns summed over the FLAG_PHASE_LOG calls since the last collect, number of calls)."""
    t, n = (C.c_double * 4)(), C.c_uint()
    _check(lib().gemmul8_b200_phase_log_collect(C.byref(t), C.byref(n)))
    return list(t), n.value


def launch_count():
    """Kernels launched by the library so far (process-wide)."""
    return lib().gemmul8_b200_launch_count()


def modulus(j):
    return lib().gemmul8_b200_modulus(j)


def version():
    return lib().gemmul8_b200_version().decode()


# ---------------------------------------------------------------------------------------------
# measurement helpers (tests / bench)
# ---------------------------------------------------------------------------------------------
def phi_matrix(rows, cols, phi, dtype, seed=123456, device="cuda"):
    """The reference's synthetic input (GEMMul8/testing/make_matrix.hpp:7-57), column-major rows x cols,
    returned as a (cols, rows) torch tensor whose memory IS the column-major matrix."""
    import torch
    out = torch.empty((cols, rows), dtype=dtype, device=device)
    rc = aux().gemmul8_aux_phi_matrix(_dtype_tag(out), rows * cols, out.data_ptr(), float(phi), seed,
                                      torch.cuda.current_stream().cuda_stream)
    if rc:
        raise Gemmul8Error(f"phi_matrix failed ({rc})")
    return out


def dd_gemm(m, n, k, A, lda, B, ldb, transA=False, transB=False, rows=None, cols=None):
    """Double-double truth (C1, C2) with C1 + C2 ~= op(A) op(B); optional row / column samples (int32 tensors)."""
    import torch
    mm = m if rows is None else rows.numel()
    nn = n if cols is None else cols.numel()
    C1 = torch.empty((nn, mm), dtype=torch.float64, device=A.device)
    C2 = torch.empty_like(C1)
    rc = aux().gemmul8_aux_dd_gemm(mm, nn, k, A.data_ptr(), lda, int(transA), B.data_ptr(), ldb, int(transB),
                                   rows.data_ptr() if rows is not None else None,
                                   cols.data_ptr() if cols is not None else None,
                                   C1.data_ptr(), C2.data_ptr(), mm, torch.cuda.current_stream().cuda_stream)
    if rc:
        raise Gemmul8Error(f"dd_gemm failed ({rc})")
    return C1, C2
