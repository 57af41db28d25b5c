"""C-ABI surface and host logic (no GPU): symbols, workSize / layout against the reference's formulas,
error behaviour, and the "no CPU fallback" rule."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported(g):
    header = open(os.path.join(ROOT, "include", "gemmul8_b200.h")).read()
    declared = set(re.findall(r"\b(gemmul8_b200_\w+)\s*\(", header))
    assert declared, "header declares nothing?"
    lib = g.lib()
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/gemmul8_b200.h but not exported"
    assert declared == set(g.EXPORTED_SYMBOLS)


def test_product_library_has_no_cublas_dependency(g):
    import subprocess
    out = subprocess.run(["ldd", g.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "cublas" not in out.lower()


# values from the reference's formula (GEMMul8/src/gemmul8.cu:27-127), SURVEY.md section 8(a2)
@pytest.mark.parametrize("args,expected", [
    ((1024, 1024, 1024, 14, 0), 48238592),
    ((16384, 16384, 16384, 14, 0), 12348096512),
    ((8192, 8192, 8192, 14, 1), 8053096448),
    ((8192, 8192, 8192, 14, 3), 6174048256),
    ((8192, 8192, 8192, 14, 2), 6174048256),
])
def test_worksize_known_values(g, args, expected):
    assert g.workSize(*args) == expected


def test_worksize_matches_oracle_sweep(g, oracle):
    rng = np.random.default_rng(0)
    for _ in range(300):
        m, n, k = (int(x) for x in rng.integers(1, 5000, 3))
        N = int(rng.integers(2, 21))
        ct = int(rng.integers(0, 4))
        assert g.workSize(m, n, k, N, ct) == oracle.worksize(m, n, k, N, ct)


def test_worksize_bad_compute_type_returns_zero(g, capfd):
    assert g.workSize(8, 8, 8, 4, 7) == 0          # reference: prints "Unknown compute type", returns 0
    assert "Unknown compute type" in capfd.readouterr().err


def test_layout_is_the_reference_carve(g):
    m, n, k, N = 1001, 777, 333, 9
    L = g.work_layout(m, n, k, N)
    lda8i, m_pad = (k + 15) // 16 * 16, (m + 3) // 4 * 4
    sizeC = (m_pad * n + 15) // 16 * 16
    assert (L.lda8i, L.m_pad, L.sizeA, L.sizeB, L.sizeC) == (lda8i, m_pad, lda8i * m_pad, lda8i * n, sizeC)
    assert L.off_A8i == 0 and L.off_B8i == N * L.sizeA and L.off_C8u == L.off_B8i + N * L.sizeB
    assert L.off_C32i == L.off_C8u + N * sizeC and L.off_sftA == L.off_C32i + 4 * sizeC
    assert L.off_sftB == L.off_sftA + 2 * ((m + 15) // 16 * 16)
    assert L.total == g.workSize(m, n, k, N)
    for off in (L.off_A8i, L.off_B8i, L.off_C8u, L.off_C32i, L.off_sftA, L.off_sftB):
        assert off % 16 == 0
    Lk = g.work_layout(m, n, k, N, g.COMPLEX_KARATSUBA_MULT)
    assert Lk.off_A8i_imag == N * Lk.sizeA and Lk.off_B8i == 2 * N * Lk.sizeA
    assert Lk.total == g.workSize(m, n, k, N, g.COMPLEX_KARATSUBA_MULT)


def test_blocked_worksize_and_plan(g):
    """Low-memory call: workspace formula, block constraints and the planner (host logic only)."""
    m = n = k = 65536
    N = 14
    full = g.workSizeBlocked(m, n, k, N, m, n)
    assert full < g.workSize(m, n, k, N)                          # no int32 product matrix in our carve
    ws = g.workSizeBlocked(m, n, k, N, 16384, 16384)
    assert ws == N * k * 32768 + N * 16384 * 16384 + 6 * (m + n) + 1024   # + the pair GEMM's per-launch claim table
    assert g.workSizeBlocked(m, n, k, N, 1000, 16384) == 0        # block sizes: multiples of 256 ...
    assert g.workSizeBlocked(1000, 900, 64, N, 1000, 900) > 0     # ... or the whole dimension
    assert g.workSizeBlocked(1000, 900, 64, N, 4096, 4096) == g.workSizeBlocked(1000, 900, 64, N, 1000, 900)
    for budget in (64 << 30, 40 << 30, 8 << 30, 1 << 30):
        mb, nb, wb = g.plan_blocks(m, n, k, N, budget)
        assert wb <= budget and wb == g.workSizeBlocked(m, n, k, N, mb, nb)
        assert mb % 256 == 0 and nb % 256 == 0 and mb >= 256 and nb >= 256
    mb, nb, wb = g.plan_blocks(m, n, k, N, 64 << 30)
    assert mb == m and -(-n // nb) <= 16                           # 64 GiB: all of A resident (no re-encoding), a handful of column blocks
    assert g.plan_blocks(3000, 2000, 500, 9, 1 << 40)[:2] == (3000, 2000)     # everything fits: one block
    rng = np.random.default_rng(1)
    for _ in range(200):
        mm, nn, kk = (int(x) for x in rng.integers(1, 20000, 3))
        NN = int(rng.integers(2, 21))
        lo = g.workSizeBlocked(mm, nn, kk, NN, 256, 256)
        budget = int(rng.integers(lo, 4 * lo + 1000))
        mb, nb, wb = g.plan_blocks(mm, nn, kk, NN, budget)
        assert lo <= wb <= budget and wb == g.workSizeBlocked(mm, nn, kk, NN, mb, nb)
        assert (mb == mm or mb % 256 == 0) and (nb == nn or nb % 256 == 0)
    with pytest.raises(g.Gemmul8Error):
        g.plan_blocks(m, n, k, N, 1 << 20)                         # below the 256 x 256 minimum


def test_moduli_and_weights_exposed(g):
    mods = [g.modulus(j) for j in range(20)]
    assert mods == [256, 255, 253, 251, 247, 241, 239, 233, 229, 227, 223, 217, 211, 199, 197, 193, 191, 181, 179, 173]
    lib = g.lib()
    assert lib.gemmul8_b200_crt_weight(2, 0, 0) == 65025.0 and lib.gemmul8_b200_crt_weight(2, 1, 0) == 256.0


def _fake_args(g, dtype="float64", ct=0, N=14):
    import torch
    t = getattr(torch, dtype)
    A = torch.zeros(4, dtype=t)
    return g.make_args(0, 0, 2, 2, 1, 1.0, A, 2, A, 1, 0.0, A, 2, N, True, torch.zeros(g.workSize(2, 2, 1, 14) or 64, dtype=torch.uint8), ct, stream=0)


def test_unsupported_compute_type_is_the_reference_error(g, capfd):
    a = _fake_args(g, ct=g.COMPLEX_KARATSUBA_MULT)           # real types accept only REAL_DEFAULT
    assert g.lib().gemmul8_b200_gemm(ctypes.byref(a)) == 1
    assert "Unsupported compute type for the argument types." in capfd.readouterr().err
    assert list(a.timers_ns) == [0.0] * 4
    c = _fake_args(g, "complex128", ct=g.REAL_DEFAULT)        # complex types accept only COMPLEX_*
    assert g.lib().gemmul8_b200_gemm(ctypes.byref(c)) == 1


def test_argument_validation(g):
    a = _fake_args(g, N=21)
    assert g.lib().gemmul8_b200_gemm(ctypes.byref(a)) == 2
    a = _fake_args(g, N=1)
    assert g.lib().gemmul8_b200_gemm(ctypes.byref(a)) == 2


def test_no_cpu_fallback(g):
    """Without a CUDA device the compute entry points must fail loudly, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    a = _fake_args(g)
    rc = g.lib().gemmul8_b200_gemm(ctypes.byref(a))
    assert rc == 3 and b"no CPU fallback" in g.lib().gemmul8_b200_last_error()
    with pytest.raises(g.Gemmul8Error):
        A = torch.zeros(4, dtype=torch.float64)
        g.gemm(0, 0, 0, 2, 2, 1, 1.0, A, 2, A, 1, 0.0, A, 2, 14, True, torch.zeros(4096, dtype=torch.uint8))


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under the package or include/ may reference it."""
    bad = []
    for base in ("mixed-gemmul8_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".inc")):
                    txt = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"liboracle|oracle\.py|import oracle|oracle/_ref|libgemmul8_ref", txt):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_options_api_without_a_device(g):
    """set_option / get_option are host state; init() needs a device and says so (no CPU fallback)."""
    assert g.get_option("gemm_pair") in (-1, 0, 1)
    old = g.get_option("pair_band")
    g.set_option("pair_band", 4)
    assert g.get_option("pair_band") == 4
    g.set_option("pair_band", old)
    assert g.get_option("strips") == 0 and g.get_option("strip_calls") == 0     # pipeline by size; nothing has run here
    with pytest.raises(Exception):
        g.set_option("strip_calls", 1)                                           # a read-only counter
    with pytest.raises(Exception):
        g.set_option("strips", 9)
    with pytest.raises(g.Gemmul8Error):
        g.set_option("no_such_option", 1)
    with pytest.raises(g.Gemmul8Error):
        g.set_option("gemm_pair", 7)
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(g.Gemmul8Error):
            g.init()


def test_multi_gpu_abi_symbols_and_row_pieces(g):
    """include/gemmul8_b200_mp.h: every declared entry point is exported by libgemmul8_b200_mp.so; the row-piece partition of
    an A panel (host logic of gemmul8_b200_pgemm) covers the block on tile boundaries."""
    from importlib import import_module
    mp = import_module("gemmul8_b200.mp")
    header = open(os.path.join(ROOT, "include", "gemmul8_b200_mp.h")).read()
    declared = set(re.findall(r"\b(gemmul8_b200_\w+)\s*\(", header))
    lib = mp.lib()
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/gemmul8_b200_mp.h but not exported"
    assert declared == set(mp.EXPORTED_SYMBOLS)
    for rows in (1, 255, 256, 257, 300, 777, 1000, 4096, 16384, 32768, 33000):
        for want in (1, 2, 4, 8):
            pcs = mp.row_pieces(rows, want)
            assert pcs[0][0] == 0 and pcs[-1][1] == rows and len(pcs) <= want
            assert all(a[1] == b[0] for a, b in zip(pcs, pcs[1:]))
            assert all(r0 % 256 == 0 and r1 > r0 for r0, r1 in pcs)
    # the product library itself stays free of NCCL (the multi-GPU layer is a separate library)
    import subprocess
    out = subprocess.run(["ldd", g.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "nccl" not in out.lower()
