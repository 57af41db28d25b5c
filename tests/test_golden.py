"""Golden vectors dumped from the UNMODIFIED reference (tests/golden/make_golden.py, run on a B200):
the CPU oracle must reproduce them (CPU test), and so must the CUDA path (GPU test)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, golden_files

NP = {"f32": np.float32, "f64": np.float64}
FILES = golden_files()


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
