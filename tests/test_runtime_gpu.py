"""Runtime behaviour of the library around the kernels: work ownership of the CTA-pair GEMM under
concurrent kernels, device-resident alpha / beta, explicit initialisation, a graph capture as the very first call."""
import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _call(g, torch, m, n, k, N, A, B, stream=None, alpha=1.0, beta=0.0, C=None, flags=0, fast=True, ct=0):
    C = torch.zeros((n, m), dtype=A.dtype, device="cuda") if C is None else C
    work = torch.zeros(g.workSize(m, n, k, N, ct), dtype=torch.uint8, device="cuda")
    g.gemm(stream, 0, 0, m, n, k, alpha, A, m, B, k, beta, C, m, N, fast, work, computeType=ct, flags=flags)
    return C, work


def test_pair_kernel_two_streams_concurrently(g):
    """the pair kernel used to derive its work from %smid alone.  With a second persistent kernel
    resident, clusters are placed wherever an SM pair frees up, possibly twice on one TPC: every cluster now claims a
    unique slot per launch, so two calls running at the same time on two streams (own workspaces) must both be right."""
    import torch
    g.init()
    m, n, k, N = 2048, 2304, 1024, 14        # 8 x 9 x 14 = 1008 items: the placed path (>= 74 pairs) is taken
    A = g.phi_matrix(m, k, 0.5, torch.float64)
    B = g.phi_matrix(k, n, 0.5, torch.float64, seed=9)
    B2 = g.phi_matrix(k, n, 0.5, torch.float64, seed=10)
    ref1, _ = _call(g, torch, m, n, k, N, A, B)
    ref2, _ = _call(g, torch, m, n, k, N, A, B2)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for it in range(6):
        keep = []
        for rep in range(3):                  # several calls back to back per stream: launches of both streams interleave
            with torch.cuda.stream(s1):
                keep.append(_call(g, torch, m, n, k, N, A, B, stream=s1))
            with torch.cuda.stream(s2):
                keep.append(_call(g, torch, m, n, k, N, A, B2, stream=s2))
        torch.cuda.synchronize()
        for i, (C, _) in enumerate(keep):
            assert torch.equal(C, ref1 if i % 2 == 0 else ref2), (it, i)


def test_pair_kernel_beside_a_resident_kernel(g):
    """The same with a foreign kernel holding SMs while the GEMM runs (what NCCL's kernels do in the multi-GPU path)."""
    import torch
    m, n, k, N = 2048, 2048, 2048, 14
    A = g.phi_matrix(m, k, 0.5, torch.float64)
    B = g.phi_matrix(k, n, 0.5, torch.float64, seed=3)
    ref, _ = _call(g, torch, m, n, k, N, A, B)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    x = torch.randn(4096, 4096, device="cuda")
    for it in range(4):
        with torch.cuda.stream(side):
            for _ in range(4):
                x = torch.tanh(x @ x) * 0.5    # cuBLAS + elementwise kernels competing for SMs
        C, _ = _call(g, torch, m, n, k, N, A, B)
        torch.cuda.synchronize()
        assert torch.equal(C, ref), it


@pytest.mark.parametrize("dt,ct", [("float64", 0), ("float32", 0), ("complex128", 3), ("complex64", 1)])
@pytest.mark.parametrize("alpha,beta", [(1.0, 0.0), (0.75, -1.5), (1.0, 1.0), (2.0, 0.0)])
def test_device_scalars_equal_host_scalars(g, dt, ct, alpha, beta):
    """GEMMUL8_FLAG_DEVICE_SCALARS: alpha / beta live in device memory and are read by the CRT kernel (the reference reads
    them on the host, GEMMul8/src/gemmul8.cu:288): same bits as the host-scalar call."""
    import torch
    dtype = getattr(torch, dt)
    m, n, k, N = 300, 260, 200, 9 if "64" in dt or dt == "complex128" else 6
    A = g.phi_matrix(m, k, 0.5, dtype)
    B = g.phi_matrix(k, n, 0.5, dtype, seed=2)
    C0 = g.phi_matrix(m, n, 1.0, dtype, seed=5)
    if dtype.is_complex:
        alpha, beta = complex(alpha, 0.25), complex(beta, -0.5) if beta not in (0.0, 1.0) else complex(beta, 0.0)
    Ch, _ = _call(g, torch, m, n, k, N, A, B, alpha=alpha, beta=beta, C=C0.clone(), ct=ct)
    da = torch.tensor([alpha], dtype=dtype, device="cuda")
    db = torch.tensor([beta], dtype=dtype, device="cuda")
    Cd, _ = _call(g, torch, m, n, k, N, A, B, alpha=da, beta=db, C=C0.clone(), ct=ct, flags=g.FLAG_DEVICE_SCALARS)
    torch.cuda.synchronize()
    assert torch.equal(torch.view_as_real(Ch) if dtype.is_complex else Ch, torch.view_as_real(Cd) if dtype.is_complex else Cd)
    # k == 0: C = beta * C, also from device memory
    Ck = C0.clone()
    g.gemm(None, 0, 0, m, n, 0, da, A, m, B, 1, db, Ck, m, N, True, torch.zeros(64, dtype=torch.uint8, device="cuda"), computeType=ct,
           flags=g.FLAG_DEVICE_SCALARS)
    torch.cuda.synchronize()
    want = C0 * beta if beta != 0 else torch.zeros_like(C0)
    assert torch.allclose(torch.view_as_real(Ck) if dtype.is_complex else Ck, torch.view_as_real(want) if dtype.is_complex else want,
                          rtol=1e-6 if "32" in dt or dt == "complex64" else 1e-14, atol=0)


def test_init_and_options(g):
    import torch
    g.init()
    g.init(0)
    with pytest.raises(g.Gemmul8Error):
        g.init(torch.cuda.device_count())
    g.set_option("gemm_pair", 0)
    assert g.get_option("gemm_pair") == 0
    g.set_option("gemm_pair", -1)


def test_graph_capture_as_the_very_first_call():
    """A fresh process whose FIRST gemm is captured into a CUDA graph (no warm-up, no init): the lazy placement probe must
    not run inside the capture; the replayed graph gives the bits of an eager call."""
    code = textwrap.dedent("""
        import sys
        sys.path.insert(0, %r)
        import torch, gemmul8_b200 as g
        m, n, k, N = 1536, 1280, 640, 14
        A = g.phi_matrix(m, k, 0.5, torch.float64); B = g.phi_matrix(k, n, 0.5, torch.float64, seed=4)
        C = torch.zeros((n, m), dtype=torch.float64, device="cuda")
        work = torch.zeros(g.workSize(m, n, k, N), dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=s):
                g.gemm(s, 0, 0, m, n, k, 1.0, A, m, B, k, 0.0, C, m, N, True, work)
            gr.replay()
        torch.cuda.synchronize()
        C1 = C.clone()
        C.zero_()
        g.gemm(None, 0, 0, m, n, k, 1.0, A, m, B, k, 0.0, C, m, N, True, work)   # eager: probes now, placed path
        torch.cuda.synchronize()
        assert torch.equal(C, C1) and C.abs().sum().item() > 0
        gr.replay(); torch.cuda.synchronize()
        assert torch.equal(C, C1)
        print("capture-first ok")
    """ % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "capture-first ok" in r.stdout, r.stdout + r.stderr
