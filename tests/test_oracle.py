"""The CPU oracle checked from first principles (exact integer arithmetic in Python), independent of
any GPU: residue encoding, the error-free product, residue reduction, CRT reconstruction, shifts."""
import math

import numpy as np
import pytest

MOD = [256, 255, 253, 251, 247, 241, 239, 233, 229, 227, 223, 217, 211, 199, 197, 193, 191, 181, 179, 173]


def phi(rng, r, c, p, dt=np.float64):
    """column-major r x c matrix as a (c, r) array, (U - 0.5) * exp(p * Z)"""
    return ((rng.random((c, r)) - 0.5) * np.exp(p * rng.standard_normal((c, r)))).astype(dt)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("strided", [0, 1])
def test_encode_is_the_symmetric_residue(oracle, dt, strided):
    rng = np.random.default_rng(1)
    nvec, length, N = 7, 37, 20
    X = phi(rng, nvec, length, 1.0, dt) if strided else phi(rng, length, nvec, 1.0, dt)
    ld = nvec if strided else length
    sft = rng.integers(-60, -10, nvec).astype(np.int16) if dt == np.float64 else rng.integers(-22, -5, nvec).astype(np.int16)
    ld8i = (length + 15) // 16 * 16
    out = oracle.encode(X, strided, nvec, length, ld, sft, N, ld8i)
    for v in range(nvec):
        for i in range(ld8i):
            for j in range(N):
                r = int(out[j, v, i])
                if i >= length:
                    assert r == 0
                    continue
                x = X[i, v] if strided else X[v, i]
                xi = int(math.trunc(math.ldexp(float(x), -int(sft[v]))))
                m = MOD[j]
                assert (r - xi) % m == 0, (v, i, j, r, xi)
                assert -m / 2 <= r <= m / 2 and -128 <= r <= 127
                if m == 256 and xi % 256 == 128:
                    assert r == -128        # +128 wraps


def test_int8_gemm_and_residue(oracle):
    rng = np.random.default_rng(2)
    A = rng.integers(-128, 128, (33, 48), dtype=np.int8)
    B = rng.integers(-128, 128, (21, 48), dtype=np.int8)
    C = oracle.int8_gemm(A, B)
    ref = (B.astype(np.int64) @ A.astype(np.int64).T)
    assert np.array_equal(C.astype(np.int64), ref)
    big = rng.integers(-2 ** 31, 2 ** 31, 5000, dtype=np.int64).astype(np.int32)
    for j in (0, 1, 7, 19):
        assert np.array_equal(oracle.residue(big, j).astype(np.int64), big.astype(np.int64) % MOD[j])


def test_int8_gemm_wraps_mod_2_32(oracle):
    k = 1 << 17
    A = np.full((1, k), -128, np.int8)
    B = np.full((1, k), -128, np.int8)
    assert int(oracle.int8_gemm(A, B)[0, 0]) == ((128 * 128 * k + 2 ** 31) % 2 ** 32) - 2 ** 31


@pytest.mark.parametrize("N,fast", [(14, 1), (14, 0), (8, 1), (7, 1), (20, 1), (4, 0)])
def test_pipeline_is_exact_up_to_truncation(oracle, N, fast):
    """C must equal 2^(sA+sB) * (A^ B^) exactly, where A^, B^ are the truncated scaled integers:
    the CRT reconstruction is error-free as long as |A^ B^| < M/2 (which the shift choice guarantees)."""
    rng = np.random.default_rng(3)
    m, n, k = 24, 20, 40
    A, B = phi(rng, m, k, 0.5), phi(rng, k, n, 0.5)
    C = np.zeros((n, m))
    r = oracle.gemm_real(0, 0, m, n, k, 1.0, A, m, B, k, 0.0, C, m, N, fast)
    M = math.prod(MOD[:N])
    for col in range(n):
        for row in range(m):
            acc = 0
            for p in range(k):
                a = int(math.trunc(math.ldexp(float(A[p, row]), -int(r.sftA[row]))))
                b = int(math.trunc(math.ldexp(float(B[col, p]), -int(r.sftB[col]))))
                acc += a * b
            assert abs(acc) < M // 2, "shift selection must keep the exact product inside the CRT range"
            e = int(r.sftA[row]) + int(r.sftB[col])
            want = math.ldexp(float(acc), e)   # float(acc) rounds once; C rounds once too
            got = C[col, row]
            if N >= 8:   # split weights: the weighted sum is error-free, only final roundings remain
                assert got == want or abs(got - want) <= abs(want) * 2 ** -50, (row, col, got, want)
            else:        # single weights (N <= 7): sum_j w_j r_j is rounded at magnitude N*255*M, as in the reference
                assert abs(got - want) <= math.ldexp(N * 255 * M * 2.0 ** -51, e), (row, col, got, want)


def test_accuracy_levels(oracle):
    """relerr against a double-double truth falls with the number of moduli like the reference's table
    (1.8e5 at N=2 ... ~1e-15 at N>=18; oz2_results_d_accuracy_*.csv:3)."""
    rng = np.random.default_rng(4)
    m = n = k = 48
    A, B = phi(rng, m, k, 0.5), phi(rng, k, n, 0.5)
    C1, C2 = oracle.dd_gemm(m, n, k, A, m, B, k)
    prev = None
    for N in (6, 10, 14, 18):
        C = np.zeros((n, m))
        oracle.gemm_real(0, 0, m, n, k, 1.0, A, m, B, k, 0.0, C, m, N, 1)
        med = np.median(np.abs((C - C1 - C2) / C1))
        assert prev is None or med < prev
        prev = med
    assert prev < 1e-15


def test_transposes_and_alpha_beta(oracle):
    rng = np.random.default_rng(5)
    m, n, k, N = 20, 12, 30, 12
    A, B = phi(rng, m, k, 0.5), phi(rng, k, n, 0.5)
    C0 = phi(rng, m, n, 0.5)
    base = np.zeros((n, m))
    oracle.gemm_real(0, 0, m, n, k, 1.0, A, m, B, k, 0.0, base, m, N, 1)
    At, Bt = np.ascontiguousarray(A.T), np.ascontiguousarray(B.T)       # column-major k x m, n x k
    Ct = np.zeros((n, m))
    oracle.gemm_real(1, 1, m, n, k, 1.0, At, k, Bt, n, 0.0, Ct, m, N, 1)
    assert np.array_equal(base, Ct)
    Cab = C0.copy()
    oracle.gemm_real(0, 0, m, n, k, 2.0, A, m, B, k, -0.5, Cab, m, N, 1)
    from fractions import Fraction
    fma = np.vectorize(lambda a, b, c: float(Fraction(a) * Fraction(b) + Fraction(c)))   # exact, rounded once
    assert np.array_equal(Cab, fma(-0.5, C0, 2.0 * base))
    Cb = C0.copy()
    oracle.gemm_real(0, 0, m, n, k, 1.0, A, m, B, k, 1.0, Cb, m, N, 1)
    assert np.array_equal(Cb, base + C0)


def test_zero_vectors_and_degenerate_shapes(oracle):
    rng = np.random.default_rng(6)
    m, n, k, N = 9, 5, 17, 14
    A, B = phi(rng, m, k, 0.5), phi(rng, k, n, 0.5)
    A[:, 3] = 0.0        # a zero row of A
    B[2, :] = 0.0        # a zero column of B
    for fast in (1, 0):
        C = np.full((n, m), 7.0)
        oracle.gemm_real(0, 0, m, n, k, 1.0, A, m, B, k, 0.0, C, m, N, fast)
        assert np.all(C[:, 3] == 0) and np.all(C[2, :] == 0) and np.isfinite(C).all()
    for shape in [(1, 1, 1), (1, 7, 3), (5, 1, 2), (3, 3, 1)]:
        mm, nn, kk = shape
        a, b = phi(rng, mm, kk, 0.5), phi(rng, kk, nn, 0.5)
        C = np.zeros((nn, mm))
        oracle.gemm_real(0, 0, mm, nn, kk, 1.0, a, mm, b, kk, 0.0, C, mm, 16, 1)
        assert np.allclose(C, b @ a, rtol=1e-12)


def test_reference_residue_chain_is_the_exact_symmetric_residue(oracle):
    """mod_8i (GEMMul8/src/scaling.hpp:215-230), restated in oracle.c, against integer arithmetic for
    magnitudes up to 2^79 (20 moduli): it always lands on the symmetric residue (the int8 wrap maps
    +128 to -128 for modulus 256).  This is what allows the CUDA encoders to use shorter exact routes."""
    import ctypes as C
    L = oracle.cpu()
    L.oracle_mod8_f.restype = C.c_int
    L.oracle_mod8_f.argtypes = [C.c_float, C.c_uint]
    L.oracle_mod8_d.restype = C.c_int
    L.oracle_mod8_d.argtypes = [C.c_double, C.c_uint]
    mods = [256, 255, 253, 251, 247, 241, 239, 233, 229, 227, 223, 217, 211, 199, 197, 193, 191, 181, 179, 173]
    rng = np.random.default_rng(1)

    def canon(a, m):
        r = a % m
        if r > m // 2:
            r -= m
        return -128 if (m == 256 and r == 128) else r

    for bits, fn, wide in ((24, L.oracle_mod8_f, False), (53, L.oracle_mod8_d, True)):
        for ebits in range(1, 80):
            for _ in range(60):
                mant = int(rng.integers(1 << (bits - 1), 1 << bits))
                e = ebits - bits
                a = mant * (2 ** e) if e >= 0 else mant >> (-e)
                a = -a if rng.random() < 0.5 else a
                x = float(a) if wide else float(np.float32(a))
                if int(x) != a:
                    continue
                for j in (0, 1, 5, 13, 19):
                    assert fn(x, j) == canon(a, mods[j]), (a, mods[j])
