"""ctypes doors to the test-only checkers (see oracle/oracle.c header: TEST INFRASTRUCTURE ONLY).

  cpu()  -> oracle/_ref/liboracle.so       CPU restatement of the reference path (numpy in / out)
  ref()  -> oracle/_ref/libgemmul8_ref.so  the unmodified reference, driven on the GPU (device pointers)

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "_ref", "liboracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libgemmul8_ref.so")
F32, F64, C32, C64 = 0, 1, 2, 3
_NP_TAG = {np.dtype(np.float32): F32, np.dtype(np.float64): F64, np.dtype(np.complex64): C32, np.dtype(np.complex128): C64}

_cpu = None
_ref = None


def build():
    """(Re)build what can be built here; the reference library needs /root/reference."""
    subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)


def cpu():
    global _cpu
    if _cpu is None:
        if not os.path.exists(ORACLE_SO):
            build()
        L = C.CDLL(ORACLE_SO)
        L.oracle_worksize.restype = C.c_size_t
        L.oracle_worksize.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint, C.c_int]
        L.oracle_num_threads.restype = C.c_int
        _cpu = L
    return _cpu


def have_ref():
    return os.path.exists(REF_SO)


def ref():
    global _ref
    if _ref is None:
        L = C.CDLL(REF_SO)
        L.gemmul8_ref_worksize.restype = C.c_size_t
        L.gemmul8_ref_worksize.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint, C.c_int]
        L.gemmul8_ref_gemm.restype = C.c_int
        L.gemmul8_ref_gemm.argtypes = [C.c_int] * 5 + [C.c_size_t] * 3 + [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                                                          C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint, C.c_int,
                                                                          C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        _ref = L
    return _ref


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def worksize(m, n, k, N, ct=0):
    return cpu().oracle_worksize(m, n, k, N, ct)


def fast_shifts(X, strided, nvec, length, ld, W, N):
    """X: numpy array holding a column-major matrix.  Returns (negated shifts int16, ambiguous uint8)."""
    sft = np.zeros(nvec, np.int16)
    amb = np.zeros(nvec, np.uint8)
    cpu().oracle_fast_shifts(C.c_int(_NP_TAG[X.dtype]), C.c_int(int(strided)), _p(X), C.c_size_t(ld), C.c_size_t(nvec),
                             C.c_size_t(length), C.c_int(W), C.c_uint(N), _p(sft), _p(amb))
    return sft, amb


def encode(X, strided, nvec, length, ld, sft_neg, N, ld8i, rows_alloc=None):
    rows_alloc = rows_alloc or nvec
    out = np.zeros((N, rows_alloc, ld8i), np.int8)
    cpu().oracle_encode(C.c_int(_NP_TAG[X.dtype]), C.c_int(int(strided)), _p(X), C.c_size_t(ld), C.c_size_t(nvec), C.c_size_t(length),
                        _p(np.ascontiguousarray(sft_neg, np.int16)), C.c_uint(N), _p(out), C.c_size_t(ld8i), C.c_size_t(rows_alloc * ld8i))
    return out


def int8_gemm(A8i, B8i):
    """A8i (m, ld8i), B8i (n, ld8i) int8 -> int32 (n, m): column-major m x n product."""
    m, ld8i = A8i.shape
    n = B8i.shape[0]
    out = np.zeros((n, m), np.int32)
    cpu().oracle_int8_gemm(C.c_size_t(m), C.c_size_t(n), C.c_size_t(ld8i), _p(np.ascontiguousarray(A8i)),
                           _p(np.ascontiguousarray(B8i)), _p(out), C.c_size_t(m))
    return out


def residue(C32i, j):
    out = np.zeros(C32i.shape, np.uint8)
    cpu().oracle_residue(C.c_size_t(C32i.size), _p(np.ascontiguousarray(C32i)), C.c_uint(j), _p(out))
    return out


class CpuResult:
    pass


def gemm_real(op_A, op_B, m, n, k, alpha, A, lda, B, ldb, beta, Cmat, ldc, N, fastmode):
    """Whole real path on the CPU.  A, B, Cmat: numpy arrays whose memory is the column-major matrix
    (Cmat is updated in place).  Returns an object with the reference-layout workspace views."""
    ws = worksize(m, n, k, N, 0)
    work = np.zeros(ws, np.uint8)
    amb_r = np.zeros(m, np.uint8)
    amb_c = np.zeros(n, np.uint8)
    rc = cpu().oracle_gemm_real(C.c_int(op_A), C.c_int(op_B), C.c_size_t(m), C.c_size_t(n), C.c_size_t(k), C.c_double(alpha),
                                C.c_int(_NP_TAG[A.dtype]), _p(A), C.c_size_t(lda), C.c_int(_NP_TAG[B.dtype]), _p(B), C.c_size_t(ldb),
                                C.c_double(beta), C.c_int(_NP_TAG[Cmat.dtype]), _p(Cmat), C.c_size_t(ldc), C.c_uint(N),
                                C.c_int(int(fastmode)), _p(work), _p(amb_r), _p(amb_c))
    if rc:
        raise RuntimeError(f"oracle_gemm_real failed ({rc})")
    r = CpuResult()
    ld8 = (k + 15) // 16 * 16
    m_pad = (m + 3) // 4 * 4
    sizeA, sizeB = ld8 * m_pad, ld8 * n
    sizeC = (m_pad * n + 15) // 16 * 16
    o = 0
    r.A8i = work[o:o + N * sizeA].view(np.int8).reshape(N, m_pad, ld8); o += N * sizeA
    r.B8i = work[o:o + N * sizeB].view(np.int8).reshape(N, n, ld8); o += N * sizeB
    r.C8u = work[o:o + N * sizeC].reshape(N, sizeC)[:, :m_pad * n].reshape(N, n, m_pad); o += N * sizeC
    o += 4 * sizeC
    r.sftA = work[o:o + 2 * m].view(np.int16); o += 2 * ((m + 15) // 16 * 16)
    r.sftB = work[o:o + 2 * n].view(np.int16)
    r.amb_rows, r.amb_cols, r.work = amb_r, amb_c, work
    return r


def gemm_complex(op_A, op_B, m, n, k, A, lda, B, ldb, Cmat, ldc, N, fastmode, ct):
    """Whole complex path on the CPU (alpha = 1, beta = 0).  Returns reference-layout workspace views
    (big matrix: A8i / B8i / C8u_real / C8u_imag; otherwise A8i_real, A8i_imag, ...)."""
    ws = worksize(m, n, k, N, ct)
    work = np.zeros(ws + 16, np.uint8)
    amb_r = np.zeros(m, np.uint8)
    amb_c = np.zeros(n, np.uint8)
    rc = cpu().oracle_gemm_complex(C.c_int(op_A), C.c_int(op_B), C.c_size_t(m), C.c_size_t(n), C.c_size_t(k),
                                   C.c_int(_NP_TAG[A.dtype]), _p(A), C.c_size_t(lda), C.c_int(_NP_TAG[B.dtype]), _p(B), C.c_size_t(ldb),
                                   C.c_int(_NP_TAG[Cmat.dtype]), _p(Cmat), C.c_size_t(ldc), C.c_uint(N), C.c_int(int(fastmode)),
                                   C.c_int(ct), _p(work), _p(amb_r), _p(amb_c))
    if rc:
        raise RuntimeError(f"oracle_gemm_complex failed ({rc})")
    r = CpuResult()
    big = ct == 1
    ld8 = ((2 * k if big else k) + 15) // 16 * 16
    m_pad = ((2 * m if big else m) + 3) // 4 * 4
    sizeA, sizeB = ld8 * m_pad, ld8 * n
    sizeC = (m_pad * n + 15) // 16 * 16
    parts = 1 if big else 2
    o = 0

    def stack8(off, rows):
        return work[off:off + N * rows * ld8].view(np.int8).reshape(N, rows, ld8)

    def stackC(off):
        return work[off:off + N * sizeC].reshape(N, sizeC)[:, :m_pad * n].reshape(N, n, m_pad)

    if big:
        r.A8i = stack8(0, m_pad); o = N * sizeA
        r.B8i = stack8(o, n); o += N * sizeB
        c = stackC(o); o += N * sizeC
        r.C8u_real, r.C8u_imag = c[:, :, :m], c[:, :, m:2 * m]
    else:
        r.A8i_real, r.A8i_imag = stack8(0, m_pad), stack8(N * sizeA, m_pad); o = 2 * N * sizeA
        r.B8i_real, r.B8i_imag = stack8(o, n), stack8(o + N * sizeB, n); o += 2 * N * sizeB
        r.C8u_real, r.C8u_imag = stackC(o)[:, :, :m], stackC(o + N * sizeC)[:, :, :m]; o += 2 * N * sizeC
    o += parts * 4 * sizeC
    r.sftA = work[o:o + 2 * m].view(np.int16); o += 2 * ((m + 15) // 16 * 16)
    r.sftB = work[o:o + 2 * n].view(np.int16)
    r.amb_rows, r.amb_cols = amb_r, amb_c
    return r


def dd_gemm(m, n, k, A, lda, B, ldb):
    """Host double-double reference GEMM (restated eval::dd::simple_gemm); returns (C1, C2) as (n, m) arrays."""
    C1 = np.zeros((n, m), np.float64)
    C2 = np.zeros((n, m), np.float64)
    cpu().oracle_dd_gemm(C.c_size_t(m), C.c_size_t(n), C.c_size_t(k), _p(A), C.c_size_t(lda), _p(B), C.c_size_t(ldb), _p(C1), _p(C2),
                         C.c_size_t(m))
    return C1, C2


def num_threads():
    return cpu().oracle_num_threads()


# ---------------------------------------------------------------------------------------------
# the unmodified reference on the GPU (torch tensors = device memory)
# ---------------------------------------------------------------------------------------------
def ref_worksize(m, n, k, N, ct=0):
    return ref().gemmul8_ref_worksize(m, n, k, N, ct)


def ref_gemm(op_A, op_B, m, n, k, alpha, A, lda, B, ldb, beta, Cmat, ldc, N, fastmode, work, ct=0):
    """Drive gemmul8::gemm of the reference; tensors are CUDA torch tensors.  Returns the 4 timers (ns)."""
    import torch
    tag = {torch.float32: F32, torch.float64: F64, torch.complex64: C32, torch.complex128: C64}
    tc = tag[Cmat.dtype]

    def scalar(v):
        if tc == F32:
            return (C.c_float * 1)(float(v))
        if tc == F64:
            return (C.c_double * 1)(float(v))
        v = complex(v)
        return ((C.c_float if tc == C32 else C.c_double) * 2)(v.real, v.imag)

    al, be = scalar(alpha), scalar(beta)
    timers = (C.c_double * 4)()
    torch.cuda.synchronize()
    rc = ref().gemmul8_ref_gemm(tag[A.dtype], tag[B.dtype], tc, op_A, op_B, m, n, k, C.addressof(al), A.data_ptr(), lda,
                                B.data_ptr(), ldb, C.addressof(be), Cmat.data_ptr(), ldc, N, int(bool(fastmode)), work.data_ptr(), ct, timers)
    torch.cuda.synchronize()
    if rc:
        raise RuntimeError(f"reference gemm failed ({rc})")
    return list(timers)
