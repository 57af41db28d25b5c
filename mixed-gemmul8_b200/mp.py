"""ctypes binding of the multi-GPU C ABI (include/gemmul8_b200_mp.h, libgemmul8_b200_mp.so): a P x Q grid of processes, one
GPU each; panels exchanged over NCCL or over copy engines (CUDA IPC).  torch.distributed is used for ONE thing here: shipping
the 128-byte bootstrap id from rank 0 to the other ranks."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
MP_PATH = os.environ.get("GEMMUL8_B200_MP_LIB") or os.path.join(_HERE, "libgemmul8_b200_mp.so")   # (override: A/B timing of builds)
EXCHANGE_NCCL, EXCHANGE_COPY = 0, 1
ID_BYTES = 128

EXPORTED_SYMBOLS = ("gemmul8_b200_mp_unique_id", "gemmul8_b200_grid_create", "gemmul8_b200_grid_create_from_comm", "gemmul8_b200_grid_destroy",
                    "gemmul8_b200_grid_coords", "gemmul8_b200_pgemm_worksize", "gemmul8_b200_pgemm", "gemmul8_b200_mp_row_pieces",
                    "gemmul8_b200_mp_last_error")


class PArgs(C.Structure):
    """gemmul8_b200_pargs"""
    _fields_ = [("m", C.c_size_t), ("n", C.c_size_t), ("k", C.c_size_t), ("alpha", C.c_void_p),
                ("a_slice", C.c_void_p), ("lda", C.c_size_t), ("b_slice", C.c_void_p), ("ldb", C.c_size_t), ("beta", C.c_void_p),
                ("c_block", C.c_void_p), ("ldc", C.c_size_t), ("num_moduli", C.c_uint), ("fastmode", C.c_int), ("work", C.c_void_p),
                ("dtype_A", C.c_int), ("dtype_B", C.c_int), ("dtype_C", C.c_int), ("compute_type", C.c_int), ("stream", C.c_void_p), ("flags", C.c_uint),
                ("timers_ns", C.c_double * 4)]


class MpError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(MP_PATH):
            raise MpError(f"{MP_PATH} is missing: run the build first")
        # the product library first (same directory; the mp library resolves its symbols against it)
        from . import lib as product_lib
        product_lib()
        L = C.CDLL(MP_PATH)
        L.gemmul8_b200_mp_unique_id.restype = C.c_int
        L.gemmul8_b200_mp_unique_id.argtypes = [C.c_void_p]
        L.gemmul8_b200_grid_create.restype = C.c_int
        L.gemmul8_b200_grid_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]
        L.gemmul8_b200_grid_destroy.restype = C.c_int
        L.gemmul8_b200_grid_destroy.argtypes = [C.c_void_p]
        L.gemmul8_b200_grid_coords.restype = C.c_int
        L.gemmul8_b200_grid_coords.argtypes = [C.c_void_p, C.POINTER(C.c_int * 4)]
        L.gemmul8_b200_pgemm_worksize.restype = C.c_size_t
        L.gemmul8_b200_pgemm_worksize.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint]
        L.gemmul8_b200_pgemm.restype = C.c_int
        L.gemmul8_b200_pgemm.argtypes = [C.c_void_p, C.POINTER(PArgs)]
        L.gemmul8_b200_mp_row_pieces.restype = C.c_int
        L.gemmul8_b200_mp_row_pieces.argtypes = [C.c_size_t, C.c_int, C.POINTER(C.c_size_t * 9)]
        L.gemmul8_b200_mp_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise MpError(f"gemmul8_b200_mp status {rc}: {lib().gemmul8_b200_mp_last_error().decode()}")


def row_pieces(rows, want):
    b = (C.c_size_t * 9)()
    n = lib().gemmul8_b200_mp_row_pieces(rows, want, C.byref(b))
    return [(b[i], b[i + 1]) for i in range(n)]


class Grid:
    """P x Q grid over the ranks of the default torch.distributed process group (rank = p * Q + q)."""

    def __init__(self, P, Q, a_panel_bytes, b_panel_bytes, exchange=EXCHANGE_COPY):
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        idbuf = (C.c_ubyte * ID_BYTES)()
        if rank == 0:
            _check(lib().gemmul8_b200_mp_unique_id(idbuf))
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.tensor(list(idbuf), dtype=torch.uint8, device=dev)
        dist.broadcast(t, 0)
        idbuf = (C.c_ubyte * ID_BYTES)(*t.cpu().tolist())
        self.handle = C.c_void_p()
        self.exchange = exchange
        _check(lib().gemmul8_b200_grid_create(idbuf, rank, world, P, Q, a_panel_bytes, b_panel_bytes, exchange, C.byref(self.handle)))
        self.P, self.Q, self.p, self.q = P, Q, rank // Q, rank % Q

    def worksize(self, m, n, k, num_moduli):
        return lib().gemmul8_b200_pgemm_worksize(self.handle, m, n, k, num_moduli)

    def pgemm(self, m, n, k, alpha, a_slice, lda, b_slice, ldb, beta, c_block, ldc, num_moduli, fastmode, work, flags=0, stream=None,
              computeType=0):
        """C_block(p, q) = alpha * A_panel(p) * B_panel(q) + beta * C_block; tensors hold column-major data (see the header)."""
        import torch
        from . import _dtype_tag, _scalar_buf
        a = PArgs()
        a.m, a.n, a.k = m, n, k
        a.dtype_A, a.dtype_B, a.dtype_C = _dtype_tag(a_slice), _dtype_tag(b_slice), _dtype_tag(c_block)
        al, be = _scalar_buf(alpha, a.dtype_C), _scalar_buf(beta, a.dtype_C)
        a.alpha, a.beta = C.addressof(al), C.addressof(be)
        a.a_slice, a.lda, a.b_slice, a.ldb, a.c_block, a.ldc = a_slice.data_ptr(), lda, b_slice.data_ptr(), ldb, c_block.data_ptr(), ldc
        a.num_moduli, a.fastmode, a.work, a.flags = num_moduli, int(bool(fastmode)), work.data_ptr(), flags
        a.compute_type = computeType
        a.stream = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        _check(lib().gemmul8_b200_pgemm(self.handle, C.byref(a)))
        return list(a.timers_ns)

    def close(self):
        if self.handle:
            lib().gemmul8_b200_grid_destroy(self.handle)
            self.handle = C.c_void_p()
