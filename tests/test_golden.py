"""Golden vectors dumped from the UNMODIFIED reference (tests/golden/make_golden.py, run on a B200):
the CPU oracle must reproduce them (CPU test), and so must the CUDA path (GPU test)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, golden_files

NP = {"f32": np.float32, "f64": np.float64}
ALL_FILES = golden_files()
CFILES = [f for f in ALL_FILES if f[0] in "cz"]          # complex cases (tests/golden/make_golden.py COMPLEX_CASES)
FILES = [f for f in ALL_FILES if f not in CFILES]
CKEYS = ("A8i", "B8i", "A8i_real", "A8i_imag", "B8i_real", "B8i_imag", "C8u_real", "C8u_imag")


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name))
    m, n, k, N, fast, opA, opB = (int(x) for x in z["meta"])
    return z, m, n, k, N, fast, opA, opB


def test_golden_fixtures_exist():
    assert len(FILES) >= 10, "run tests/golden/make_golden.py on a GPU box and commit the .npz files"


@pytest.mark.parametrize("name", FILES)
def test_oracle_reproduces_reference(oracle, name):
    z, m, n, k, N, fast, opA, opB = load(name)
    A, B, C = z["A"], z["B"], z["C0"].copy()
    lda = A.shape[1]
    ldb = B.shape[1]
    r = oracle.gemm_real(opA, opB, m, n, k, float(z["alpha"]), A, lda, B, ldb, float(z["beta"]), C, m, N, fast)
    # shifts: identical except where the oracle flags its log2f as too close to an integer boundary
    bad_rows = (r.sftA != z["sftA"]) & (r.amb_rows == 0)
    bad_cols = (r.sftB != z["sftB"]) & (r.amb_cols == 0)
    assert not bad_rows.any() and not bad_cols.any()
    if (r.sftA == z["sftA"]).all() and (r.sftB == z["sftB"]).all():
        assert np.array_equal(r.A8i[:, :m], z["A8i"])
        assert np.array_equal(r.B8i, z["B8i"])
        assert np.array_equal(r.C8u[:, :, :m], z["C8u"])
        assert np.array_equal(C, z["C"]), "final C must be bit-identical to the reference"
    else:   # a flagged shift differs: C still agrees to the emulation's accuracy
        assert np.allclose(C, z["C"], rtol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("name", FILES)
def test_cuda_path_reproduces_reference(g, name):
    import torch
    z, m, n, k, N, fast, opA, opB = load(name)
    A, B = torch.from_numpy(z["A"]).cuda(), torch.from_numpy(z["B"]).cuda()
    C = torch.from_numpy(z["C0"].copy()).cuda()
    work = torch.zeros(g.workSize(m, n, k, N), dtype=torch.uint8, device="cuda")
    g.gemm(None, opA, opB, m, n, k, float(z["alpha"]), A, A.shape[1], B, B.shape[1], float(z["beta"]), C, m, N, bool(fast), work)
    torch.cuda.synchronize()
    v = g.work_views(work, g.work_layout(m, n, k, N), N, m, n)
    assert np.array_equal(v["sftA"].cpu().numpy(), z["sftA"]) and np.array_equal(v["sftB"].cpu().numpy(), z["sftB"])
    assert np.array_equal(v["A8i"][:, :m].cpu().numpy(), z["A8i"])
    assert np.array_equal(v["B8i"].cpu().numpy(), z["B8i"])
    assert np.array_equal(v["C8u"][:, :, :m].cpu().numpy(), z["C8u"])
    assert np.array_equal(C.cpu().numpy(), z["C"])


def load_c(name):
    z = np.load(os.path.join(GOLDEN_DIR, name))
    m, n, k, N, fast, opA, opB, ct = (int(x) for x in z["meta"])
    return z, m, n, k, N, fast, opA, opB, ct


def test_complex_golden_fixtures_exist():
    assert len(CFILES) >= 8


@pytest.mark.parametrize("name", CFILES)
def test_oracle_reproduces_reference_complex(oracle, name):
    z, m, n, k, N, fast, opA, opB, ct = load_c(name)
    A, B = z["A"], z["B"]
    C = np.zeros_like(z["C"])
    r = oracle.gemm_complex(opA, opB, m, n, k, A, A.shape[1], B, B.shape[1], C, m, N, fast, ct)
    assert not ((r.sftA != z["sftA"]) & (r.amb_rows == 0)).any()
    assert not ((r.sftB != z["sftB"]) & (r.amb_cols == 0)).any()
    if (r.sftA == z["sftA"]).all() and (r.sftB == z["sftB"]).all():
        for key in CKEYS:
            if key in z.files:
                got = getattr(r, key)
                got = got[:, :2 * m] if key == "A8i" else got[:, :m] if key.startswith("A8i_") else got
                assert np.array_equal(got, z[key]), key
        assert np.array_equal(C.view(C.real.dtype), z["C"].view(C.real.dtype)), "final C must be bit-identical to the reference"
    else:
        assert np.allclose(C, z["C"], rtol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CFILES)
def test_cuda_path_reproduces_reference_complex(g, name):
    import torch
    z, m, n, k, N, fast, opA, opB, ct = load_c(name)
    A, B = torch.from_numpy(z["A"]).cuda(), torch.from_numpy(z["B"]).cuda()
    C = torch.zeros((n, m), dtype=torch.from_numpy(z["C"]).dtype, device="cuda")
    work = torch.zeros(g.workSize(m, n, k, N, ct), dtype=torch.uint8, device="cuda")
    g.gemm(None, opA, opB, m, n, k, 1.0, A, A.shape[1], B, B.shape[1], 0.0, C, m, N, bool(fast), work, computeType=ct)
    torch.cuda.synchronize()
    v = g.work_views_complex(work, g.work_layout(m, n, k, N, ct), N, m, n, k, ct)
    assert np.array_equal(v["sftA"].cpu().numpy(), z["sftA"]) and np.array_equal(v["sftB"].cpu().numpy(), z["sftB"])
    for key in CKEYS:
        if key in z.files:
            got = v[key]
            got = got[:, :2 * m] if key == "A8i" else got[:, :m] if key.startswith("A8i_") else got
            assert np.array_equal(got.cpu().numpy(), z[key]), key
    Ch = C.cpu().numpy()
    assert np.array_equal(Ch.view(Ch.real.dtype), z["C"].view(Ch.real.dtype))
