"""Build the in-tree shared libraries with nvcc for sm_100a (cross-compiles without a GPU).

  libgemmul8_b200.so        product: C ABI (include/gemmul8_b200.h), kernels, no cuBLAS dependency
  libgemmul8_b200_blas.so   LD_PRELOAD interposer: cublasDgemm / Sgemm / Zgemm / Cgemm / GemmEx -> the product library
  libgemmul8_b200_mp.so     multi-GPU layer (include/gemmul8_b200_mp.h): 2-D block decomposition, panel exchange over NCCL or
                            copy engines; links NCCL, not part of the single-GPU product library
  libgemmul8_b200_aux.so    measurement helpers used by tests / bench only (phi-matrix generator,
                            double-double truth GEMM, cuBLAS native baselines)
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgemmul8_b200.so")
AUX = os.path.join(HERE, "libgemmul8_b200_aux.so")
BLAS = os.path.join(HERE, "libgemmul8_b200_blas.so")
MP = os.path.join(HERE, "libgemmul8_b200_mp.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return r.stdout


def build(force=False, verbose=False):
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "gemmul8_b200.h"), os.path.join(HERE, "..", "include", "gemmul8.hpp")]
    core = [os.path.join(CSRC, f) for f in ("oz_api.cu", "oz_scale.cu", "oz_gemm.cu", "oz_gemm_crt.cu", "oz_crt.cu", "oz_complex.cu", "oz_cxx_api.cu")
            if os.path.exists(os.path.join(CSRC, f))]
    out = ""
    if force or _stale(LIB, deps):
        extra = ["-Xptxas", "-v"] if verbose else []
        out += _run([nvcc(), *ARCH, *COMMON, *extra, "-shared", "-o", LIB, *core, "-ldl"])
    blas_src = os.path.join(CSRC, "oz_interpose.cpp")
    if os.path.exists(blas_src) and (force or _stale(BLAS, [blas_src, LIB])):
        # host code only (g++); resolved against the product library at load time, cuBLAS itself through RTLD_NEXT
        cuda = os.path.dirname(os.path.dirname(nvcc()))
        out += _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", os.path.join(cuda, "include"), "-o", BLAS, blas_src,
                     "-L", HERE, "-lgemmul8_b200", "-L", os.path.join(cuda, "lib64"), "-lcudart", "-ldl", "-Wl,-rpath,$ORIGIN"])
    mp_src = os.path.join(CSRC, "oz_mp.cpp")
    if os.path.exists(mp_src) and (force or _stale(MP, [mp_src, LIB, os.path.join(HERE, "..", "include", "gemmul8_b200_mp.h")])):
        # host code only: panel exchange (NCCL or copy engines over CUDA IPC) around the block-wise entry of the product library
        cuda = os.path.dirname(os.path.dirname(nvcc()))
        out += _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", os.path.join(cuda, "include"), "-o", MP, mp_src,
                     "-L", HERE, "-lgemmul8_b200", "-L", os.path.join(cuda, "lib64"), "-lcudart", "-lnccl", "-ldl", "-Wl,-rpath,$ORIGIN"])
    aux_src = os.path.join(CSRC, "oz_aux.cu")
    if os.path.exists(aux_src) and (force or _stale(AUX, [aux_src])):
        out += _run([nvcc(), *ARCH, *COMMON, "-shared", "-o", AUX, aux_src, "-lcublas"])
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
