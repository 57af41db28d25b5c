"""Complex types (cuFloatComplex / cuDoubleComplex and the mixed ones) in the three computeType_t
modes, against the unmodified reference library on the same inputs: shifts, int8 slices and
per-modulus residues bit for bit, C bit for bit where the reference itself is sound.

Known reference defect: inverse_scaling_kara launches nothing unless the weights
are single-double and (alpha, beta) = (1, 0) -- i.e. cuDoubleComplex output with N >= 8 under
CLASSIC / KARATSUBA leaves C untouched.  There C is checked against the reference's own
BIG_MATRIX_ENCODE result (the three modes produce the same residues, hence the same C)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BIG, CLASSIC, KARA = 1, 2, 3


def torch_():
    import torch
    return torch


def operands(g, m, n, k, opA, opB, dtA, dtB, phi=0.5):
    rA, cA = (m, k) if opA == 0 else (k, m)
    rB, cB = (k, n) if opB == 0 else (n, k)
    return g.phi_matrix(rA, cA, phi, dtA, seed=123456), g.phi_matrix(rB, cB, phi, dtB, seed=4242)


def run(fn_gemm, ws, g, m, n, k, N, fast, A, B, opA, opB, dtC, ct, alpha=1.0, beta=0.0, C0=None):
    torch = torch_()
    C = torch.zeros((n, m), dtype=dtC, device="cuda") if C0 is None else C0.clone()
    work = torch.zeros(ws, dtype=torch.uint8, device="cuda")
    fn_gemm(opA, opB, m, n, k, alpha, A, A.shape[1], B, B.shape[1], beta, C, m, N, fast, work, ct)
    torch.cuda.synchronize()
    return C, g.work_views_complex(work, g.work_layout(m, n, k, N, ct), N, m, n, k, ct)


def ours(g):
    return lambda opA, opB, m, n, k, al, A, lda, B, ldb, be, C, ldc, N, fast, work, ct: g.gemm(
        None, opA, opB, m, n, k, al, A, lda, B, ldb, be, C, ldc, N, fast, work, computeType=ct)


CASES = [
    # m, n, k, N, fast, opA, opB, dtA, dtB, dtC, ct
    (96, 80, 112, 14, 1, 0, 0, "complex128", "complex128", "complex128", BIG),
    (96, 80, 112, 14, 1, 0, 0, "complex128", "complex128", "complex128", KARA),
    (96, 80, 112, 14, 1, 0, 0, "complex128", "complex128", "complex128", CLASSIC),
    (70, 52, 100, 7, 1, 1, 2, "complex128", "complex128", "complex128", BIG),       # T and C, k % 16 != 0
    (70, 52, 100, 7, 1, 2, 1, "complex128", "complex128", "complex128", KARA),
    (70, 52, 101, 6, 1, 2, 2, "complex64", "complex64", "complex64", CLASSIC),      # odd k: unaligned second half
    (45, 33, 76, 15, 1, 0, 1, "complex64", "complex64", "complex64", BIG),         # (the reference needs k % 4 == 0 here)
    (129, 257, 130, 6, 1, 0, 0, "complex64", "complex64", "complex64", KARA),
    (64, 48, 64, 12, 1, 0, 0, "complex128", "complex64", "complex128", BIG),        # mixed ZC -> Z
    (64, 48, 64, 10, 1, 0, 0, "complex64", "complex128", "complex128", KARA),
    (64, 48, 64, 6, 1, 0, 0, "complex128", "complex64", "complex64", CLASSIC),
    (64, 48, 64, 6, 1, 1, 0, "complex64", "complex128", "complex64", BIG),
    (96, 80, 112, 14, 0, 0, 0, "complex128", "complex128", "complex128", BIG),      # accurate mode
    (96, 80, 112, 15, 0, 0, 0, "complex64", "complex64", "complex64", KARA),        # one_accuracy_complex.cu's setting
    # (accurate big-matrix with a transposed / conjugated A is not comparable: the reference takes the maxima for
    #  A's shifts from COLUMNS of the bound product there, scaling.hpp:3234 -> :2769-2783, and offsets the lower block
    #  row of its bound matrix by n instead of m, :3201; covered by test_accurate_bigmatrix_transposed_a below)
    (60, 44, 96, 9, 0, 1, 0, "complex128", "complex128", "complex128", KARA),       # accurate, transposed A, separate stacks
    (70, 52, 100, 8, 0, 0, 2, "complex128", "complex128", "complex128", BIG),
    (70, 52, 100, 8, 0, 0, 2, "complex128", "complex128", "complex128", CLASSIC),
    (1, 1, 1, 14, 1, 0, 0, "complex128", "complex128", "complex128", KARA),
    (5, 300, 20, 9, 1, 0, 0, "complex128", "complex128", "complex128", BIG),
    (1000, 700, 1300, 14, 1, 0, 0, "complex128", "complex128", "complex128", KARA),
    (1000, 700, 1300, 14, 1, 0, 0, "complex128", "complex128", "complex128", BIG),
]


@pytest.mark.parametrize("m,n,k,N,fast,opA,opB,dtA,dtB,dtC,ct", CASES)
def test_complex_against_unmodified_reference(g, oracle, m, n, k, N, fast, opA, opB, dtA, dtB, dtC, ct):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/libgemmul8_ref.so not built")
    torch = torch_()
    dA, dB, dC = getattr(torch, dtA), getattr(torch, dtB), getattr(torch, dtC)
    A, B = operands(g, m, n, k, opA, opB, dA, dB)
    ws = g.workSize(m, n, k, N, ct)
    assert ws == oracle.ref_worksize(m, n, k, N, ct)
    C, v = run(ours(g), ws, g, m, n, k, N, fast, A, B, opA, opB, dC, ct)
    Cr, rv = run(oracle.ref_gemm, ws, g, m, n, k, N, fast, A, B, opA, opB, dC, ct)
    assert torch.equal(v["sftA"], rv["sftA"]) and torch.equal(v["sftB"], rv["sftB"])
    if ct == BIG:
        assert torch.equal(v["A8i"][:, :2 * m], rv["A8i"][:, :2 * m])
        assert torch.equal(v["B8i"], rv["B8i"])
    else:
        # (KARATSUBA leaves (Ar + Ai) mod m_j in A8i_real, in the reference and here)
        for key in ("A8i_real", "A8i_imag"):
            assert torch.equal(v[key][:, :m], rv[key][:, :m]), key
        for key in ("B8i_real", "B8i_imag"):
            assert torch.equal(v[key], rv[key]), key
    assert torch.equal(v["C8u_real"], rv["C8u_real"]) and torch.equal(v["C8u_imag"], rv["C8u_imag"])
    ref_writes_c = ct == BIG or N <= 7 or dC == torch.complex64
    if not ref_writes_c:
        assert (Cr == 0).all()          # the defect: C untouched
        if not fast and opA != 0:
            # ... and its big-matrix accurate mode is unusable for a transposed A (shifts from the wrong maxima: the
            # result is off by percents), so the truth here is the native product
            MA, MB = A.T, B.T
            OA = MA if opA == 0 else MA.T if opA == 1 else MA.conj().T
            OB = MB if opB == 0 else MB.T if opB == 1 else MB.conj().T
            truth = (OA @ OB).T
            assert ((C - truth).abs() / truth.abs()).max().item() < (1e-9 if N >= 12 else 1e-5)
            return
        wsb = g.workSize(m, n, k, N, BIG)
        Cr, _ = run(oracle.ref_gemm, wsb, g, m, n, k, N, fast, A, B, opA, opB, dC, BIG)
    assert torch.equal(torch.view_as_real(C), torch.view_as_real(Cr))


@pytest.mark.parametrize("ct", [BIG, CLASSIC, KARA])
@pytest.mark.parametrize("alpha,beta", [(1.0, 1.0), (0.75 - 0.5j, -1.5 + 2j), (2.0 + 1j, 0.0)])
def test_complex_alpha_beta(g, ct, alpha, beta):
    """C = alpha*AB + beta*C in complex arithmetic (the reference ignores alpha/beta under CLASSIC / KARATSUBA)."""
    torch = torch_()
    m, n, k, N = 120, 90, 150, 14
    A, B = operands(g, m, n, k, 0, 0, torch.complex128, torch.complex128)
    C0 = g.phi_matrix(m, n, 1.0, torch.complex128, seed=5)
    ws = g.workSize(m, n, k, N, ct)
    P, _ = run(ours(g), ws, g, m, n, k, N, True, A, B, 0, 0, torch.complex128, ct)
    C, _ = run(ours(g), ws, g, m, n, k, N, True, A, B, 0, 0, torch.complex128, ct, alpha=alpha, beta=beta, C0=C0)
    want = alpha * P + beta * C0
    assert torch.allclose(C, want, rtol=1e-14, atol=0)     # a few ulp: fma vs separate rounding


def test_accurate_bigmatrix_transposed_a(g):
    """Accurate mode, big matrix, op_A = T / C: shifts must equal those of the separate-stack modes (same bound
    product), and the result must be as accurate as theirs."""
    torch = torch_()
    m, n, k, N = 70, 52, 100, 10
    for opA in (1, 2):
        A, B = operands(g, m, n, k, opA, 2, torch.complex128, torch.complex128)
        Cb, vb = run(ours(g), g.workSize(m, n, k, N, BIG), g, m, n, k, N, False, A, B, opA, 2, torch.complex128, BIG)
        Ck, vk = run(ours(g), g.workSize(m, n, k, N, KARA), g, m, n, k, N, False, A, B, opA, 2, torch.complex128, KARA)
        assert torch.equal(vb["sftA"], vk["sftA"]) and torch.equal(vb["sftB"], vk["sftB"])
        assert torch.equal(torch.view_as_real(Cb), torch.view_as_real(Ck))
        # column-major stored X (r x c) appears as a torch tensor of shape (c, r): math matrix = tensor.T
        MA = A.T                                     # math A (k x m)
        MB = B.T                                     # math B (n x k)
        OA = MA.T if opA == 1 else MA.conj().T       # op(A): m x k
        OB = MB.conj().T                             # op(B) = B^H: k x n
        truth = OA @ OB                              # m x n
        err = ((Cb.T - truth).abs() / truth.abs()).max().item()
        assert err < 1e-7, err                       # 10 moduli: ~2e-9 (the reference's own accuracy at N = 10)


@pytest.mark.parametrize("ct", [BIG, CLASSIC, KARA])
def test_complex_accuracy_against_native_zgemm(g, ct):
    torch = torch_()
    m, n, k, N = 256, 192, 512, 14
    A, B = operands(g, m, n, k, 0, 0, torch.complex128, torch.complex128)
    ws = g.workSize(m, n, k, N, ct)
    C, _ = run(ours(g), ws, g, m, n, k, N, True, A, B, 0, 0, torch.complex128, ct)
    want = (B @ A)      # row-major view of the column-major product A*B
    err = ((C - want).abs() / want.abs()).max().item()
    assert err < 1e-9, err


@pytest.mark.parametrize("k", [444, 445, 77, 3])
def test_three_modes_agree_bit_for_bit(g, k):
    """Also covers big-matrix calls with k % 4 != 0, where the reference itself faults (its char4 stores
    at byte offset k of a row are misaligned, scaling.hpp:797): the imaginary half is written bytewise here."""
    torch = torch_()
    m, n, N = 333, 222, 13
    A, B = operands(g, m, n, k, 2, 1, torch.complex128, torch.complex128)
    outs = []
    for ct in (BIG, CLASSIC, KARA):
        C, v = run(ours(g), g.workSize(m, n, k, N, ct), g, m, n, k, N, True, A, B, 2, 1, torch.complex128, ct)
        outs.append((C, v["C8u_real"].clone(), v["C8u_imag"].clone()))
    for C, cr, ci in outs[1:]:
        assert torch.equal(cr, outs[0][1]) and torch.equal(ci, outs[0][2])
        assert torch.equal(torch.view_as_real(C), torch.view_as_real(outs[0][0]))


def test_real_types_reject_complex_compute_type_and_vice_versa(g, capfd):
    torch = torch_()
    A = g.phi_matrix(8, 8, 0.5, torch.complex128)
    C = torch.full((8, 8), 5.0 + 0j, dtype=torch.complex128, device="cuda")
    work = torch.zeros(g.workSize(8, 8, 8, 4, KARA), dtype=torch.uint8, device="cuda")
    t = g.gemm(None, 0, 0, 8, 8, 8, 1.0, A, 8, A, 8, 0.0, C, 8, 4, True, work, computeType=g.REAL_DEFAULT)
    assert t == [0.0] * 4 and (C == 5.0).all()
    assert "Unsupported compute type" in capfd.readouterr().err


def test_no_writes_outside_workspace_and_c(g, monkeypatch):
    """compute-sanitizer is not available on this pool: guard bands of 4 KiB around `work` and around C, filled
    with a sentinel, must survive every kernel family (real fast / accurate, the three complex modes, ragged
    shapes, the opt-in CTA-pair and fused-CRT kernels)."""
    torch = torch_()
    G = 4096

    def guarded(nbytes):
        buf = torch.full((nbytes + 2 * G + 64,), 0xA5, dtype=torch.uint8, device="cuda")
        off = (-buf.data_ptr()) % 16 + G          # keep the 16-byte alignment the API asks for
        off += (-(buf.data_ptr() + off)) % 16
        return buf, buf[off:off + nbytes], off

    def check(buf, off, nbytes, what):
        assert (buf[:off] == 0xA5).all() and (buf[off + nbytes:] == 0xA5).all(), what

    cases = [(301, 203, 262, 14, True, 0, 0, torch.float64, 0, 0), (129, 257, 131, 20, False, 1, 1, torch.float64, 0, 0),
             (77, 45, 33, 6, True, 0, 1, torch.float32, 0, 0), (640, 520, 300, 14, True, 0, 0, torch.float64, 0, g.FLAG_FUSED_CRT)]
    for ct in (BIG, CLASSIC, KARA):
        cases += [(131, 70, 101, 9, True, 0, 2, torch.complex128, ct, 0), (70, 53, 100, 8, False, 1, 0, torch.complex64, ct, 0)]
    for pair in ("0", "1"):
        g.set_option("gemm_pair", int(pair))
        for (m, n, k, N, fast, opA, opB, dt, ct, flags) in cases:
            A, B = operands(g, m, n, k, opA, opB, dt, dt)
            ws = g.workSize(m, n, k, N, ct)
            wbuf, work, woff = guarded(ws)
            work.zero_()
            es = torch.empty(0, dtype=dt).element_size()
            cbuf, cbytes, coff = guarded(m * n * es)
            C = cbytes.view(dt).view(n, m)
            C.zero_()
            g.gemm(None, opA, opB, m, n, k, 1.0, A, A.shape[1], B, B.shape[1], 0.0, C, m, N, fast, work, computeType=ct, flags=flags)
            torch.cuda.synchronize()
            check(wbuf, woff, ws, ("work", m, n, k, N, ct, pair))
            check(cbuf, coff, m * n * es, ("C", m, n, k, N, ct, pair))
            assert torch.isfinite(torch.view_as_real(C) if C.is_complex() else C).all()
    g.set_option("gemm_pair", -1)


@pytest.mark.parametrize("ct", [BIG, CLASSIC, KARA])
@pytest.mark.parametrize("opA,opB", [(0, 0), (2, 1)])
def test_complex_low_memory_call_equals_gemm(g, ct, opA, opB):
    """gemmul8_b200_gemm_blocked on complex types (fast mode): C block by block in a workspace of workSize(block) bytes, same
    bits as gemm() (fast-mode shifts depend on the row / column alone), nothing written beyond the block workspace."""
    torch = torch_()
    m, n, k, N = 700, 600, 320, 14
    A, B = operands(g, m, n, k, opA, opB, torch.complex128, torch.complex128)
    C0 = g.phi_matrix(m, n, 1.0, torch.complex128, seed=5)
    ws = g.workSize(m, n, k, N, ct)
    Cf, _ = run(ours(g), ws, g, m, n, k, N, True, A, B, opA, opB, torch.complex128, ct, alpha=0.5 + 1j, beta=-2.0, C0=C0)
    mb, nb = 256, 512
    wsb = g.workSizeBlockedComplex(m, n, k, N, ct, mb, nb)
    assert 0 < wsb < ws and wsb == g.workSize(mb, nb, k, N, ct)
    buf = torch.full((wsb + 4096,), 0xA5, dtype=torch.uint8, device="cuda")
    Cb = C0.clone()
    g.gemm_blocked(None, opA, opB, m, n, k, 0.5 + 1j, A, A.shape[1], B, B.shape[1], -2.0, Cb, m, N, True, buf[:wsb], mb, nb, computeType=ct)
    torch.cuda.synchronize()
    assert torch.equal(torch.view_as_real(Cb), torch.view_as_real(Cf))
    assert bool((buf[wsb:] == 0xA5).all())
    with pytest.raises(g.Gemmul8Error):          # accurate mode is not blocked for complex types
        g.gemm_blocked(None, opA, opB, m, n, k, 1.0, A, A.shape[1], B, B.shape[1], 0.0, Cb, m, N, False, buf[:wsb], mb, nb, computeType=ct)
