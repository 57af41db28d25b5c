"""The cuBLAS interposer: an unmodified cuBLAS application preloaded with
libgemmul8_b200_blas.so gets its large D/ZGEMM calls emulated and its small ones untouched."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "mixed-gemmul8_b200")
BLAS = os.path.join(LIBDIR, "libgemmul8_b200_blas.so")
SRC = os.path.join(ROOT, "tests", "cxx", "blas_app.cu")
APP = os.path.join(ROOT, "tests", "cxx", "blas_app")


def build_app():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if os.path.exists(APP) and os.path.getmtime(APP) > os.path.getmtime(SRC):
        return
    subprocess.run([nvcc, "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-o", APP, SRC, "-lcublas"], check=True)


def test_interposer_exports_the_cublas_entry_points():
    out = subprocess.run(["nm", "-D", "--defined-only", BLAS], check=True, capture_output=True, text=True).stdout
    for sym in ("cublasDgemm_v2", "cublasSgemm_v2", "cublasZgemm_v2", "cublasCgemm_v2", "cublasGemmEx"):
        assert f" T {sym}" in out
    build_app()          # the application itself never mentions the library


def parse(text):
    vals = {}
    for line in text.splitlines():
        f = line.split()
        if f and f[0] in "DZTP":
            vals[(f[0], f[1])] = [float(x) for x in f[2:]]
    return vals


@pytest.mark.gpu
def test_unmodified_cublas_application_is_emulated():
    build_app()
    plain = subprocess.run([APP], capture_output=True, text=True, check=True)
    env = dict(os.environ, LD_PRELOAD=BLAS, GEMMUL8_VERBOSE="1", GEMMUL8_MIN_MNK=str(2 ** 24))
    pre = subprocess.run([APP], capture_output=True, text=True, env=env, check=True)
    assert "cuda status 0" in plain.stdout and "cuda status 0" in pre.stdout
    a, b = parse(plain.stdout), parse(pre.stdout)
    assert a.keys() == b.keys() and len(a) > 10
    for key in a:
        for x, y in zip(a[key], b[key]):
            assert abs(x - y) <= 1e-10 * max(1.0, abs(x)), (key, x, y)       # 14 moduli: ~1e-13 relative
    lines = [l for l in pre.stderr.splitlines() if l.startswith("gemmul8_b200_blas:")]
    # the 8^3 call stayed with cuBLAS; the call in CUBLAS_POINTER_MODE_DEVICE is emulated too (alpha / beta read on the device)
    assert len(lines) == 3 and "1024 x 768 x 2048" in lines[0] and "512 x 384 x 1024" in lines[1] and "768 x 1024 x 2048" in lines[2]
    assert any(k[0] == "P" for k in a)
    assert a[("T", "0")] == b[("T", "0")] == [16.0]


@pytest.mark.gpu
def test_interposer_takes_the_low_memory_call_when_the_workspace_is_capped():
    """GEMMUL8_MAX_WORK_MB below workSize(): the DGEMM is emulated block by block (same bits as the plain emulation),
    the ZGEMM -- no low-memory call for complex types -- stays with cuBLAS."""
    build_app()
    base = dict(os.environ, LD_PRELOAD=BLAS, GEMMUL8_VERBOSE="1", GEMMUL8_MIN_MNK=str(2 ** 24))
    full = subprocess.run([APP], capture_output=True, text=True, env=base, check=True)
    capped = subprocess.run([APP], capture_output=True, text=True, env=dict(base, GEMMUL8_MAX_WORK_MB="24"), check=True)
    a, b = parse(full.stdout), parse(capped.stdout)
    assert [v for k, v in a.items() if k[0] in "DP"] == [v for k, v in b.items() if k[0] in "DP"]      # bit-identical
    lines = [l for l in capped.stderr.splitlines() if l.startswith("gemmul8_b200_blas:")]
    assert len(lines) == 2 and "1024 x 768 x 2048" in lines[0] and all("low-memory blocks" in l for l in lines)
