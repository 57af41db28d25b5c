// libgemmul8_b200_blas.so -- cuBLAS interposer: run an UNMODIFIED application's cublasDgemm / Sgemm /
// Zgemm / Cgemm / cublasGemmEx calls through the Ozaki-II emulation.
//
//     LD_PRELOAD=<repo>/mixed-gemmul8_b200/libgemmul8_b200_blas.so  ./hpl_or_any_cublas_app
//
// This is the "caller layer" next to the hot path; the reference tree ships the
// same idea for its comparison library (ozIMMU_EF/src/cublas.cu:135-310: hijacked cublasGemmEx /
// cublasDgemm with an environment-variable mode switch).  Nothing here computes: a call is either
// forwarded to gemmul8_b200_gemm (include/gemmul8_b200.h) on the handle's stream, or passed on to the
// real cuBLAS (small problems, k > 2^17, unsupported types).  CUBLAS_POINTER_MODE_DEVICE is honoured: alpha / beta are then
// handed through as device pointers (GEMMUL8_FLAG_DEVICE_SCALARS) and read by the CRT kernel.
//
// Environment:
//   GEMMUL8_NUM_MODULI_D / _S   moduli for fp64 / fp32 results (default 14 / 6; the reference's headline settings)
//   GEMMUL8_FASTMODE            1 (default) fast mode, 0 accurate mode
//   GEMMUL8_COMPLEX             karatsuba (default) | bigmatrix | classic
//   GEMMUL8_MIN_MNK             emulate only if m*n*k >= this (default 2^27: below that launch latency dominates)
//   GEMMUL8_VERBOSE             1: one line on stderr per intercepted call
#include "../../include/gemmul8_b200.h"

#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace {

template <typename F> F next_symbol(const char *name) {
    void *p = dlsym(RTLD_NEXT, name);
    if (!p) { fprintf(stderr, "gemmul8_b200_blas: %s not found behind the interposer (is cuBLAS linked?)\n", name); abort(); }
    return reinterpret_cast<F>(p);
}

struct Config {
    unsigned moduli_d = 14, moduli_s = 6;
    int fastmode = 1, complex_type = GEMMUL8_COMPLEX_KARATSUBA_MULT, verbose = 0;
    double min_mnk = 134217728.0;
    size_t max_work = 0;   // GEMMUL8_MAX_WORK_MB: cap of the cached workspace (0 = none); larger problems take the low-memory call
    Config() {
        if (const char *e = getenv("GEMMUL8_MAX_WORK_MB")) max_work = (size_t)atoll(e) << 20;
        if (const char *e = getenv("GEMMUL8_NUM_MODULI_D")) moduli_d = (unsigned)atoi(e);
        if (const char *e = getenv("GEMMUL8_NUM_MODULI_S")) moduli_s = (unsigned)atoi(e);
        if (const char *e = getenv("GEMMUL8_FASTMODE")) fastmode = atoi(e) != 0;
        if (const char *e = getenv("GEMMUL8_VERBOSE")) verbose = atoi(e);
        if (const char *e = getenv("GEMMUL8_MIN_MNK")) min_mnk = atof(e);
        if (const char *e = getenv("GEMMUL8_COMPLEX")) {
            if (!strcmp(e, "bigmatrix")) complex_type = GEMMUL8_COMPLEX_BIG_MATRIX_ENCODE;
            else if (!strcmp(e, "classic")) complex_type = GEMMUL8_COMPLEX_CLASSIC_MULT;
        }
        if (moduli_d < 2 || moduli_d > 20) moduli_d = 14;
        if (moduli_s < 2 || moduli_s > 20) moduli_s = 6;
    }
};
const Config &config() { static Config c; return c; }

// one growing workspace PER DEVICE, each with its own mutex and its own event (created on that device): calls on one
// stream are ordered by the stream, a second stream using the buffer is ordered behind the previous call by the event
struct Workspace {
    std::mutex mu;
    void *ptr = nullptr;
    size_t bytes = 0;
    cudaEvent_t last = nullptr;
    bool recorded = false;
};
constexpr int kMaxDevices = 64;
Workspace *workspace(int dev) {
    static Workspace w[kMaxDevices];
    return (dev >= 0 && dev < kMaxDevices) ? &w[dev] : nullptr;
}

std::atomic<unsigned long long> g_intercepted{0};

int op_tag(cublasOperation_t op) { return op == CUBLAS_OP_N ? GEMMUL8_OP_N : op == CUBLAS_OP_T ? GEMMUL8_OP_T : GEMMUL8_OP_C; }

// returns true when the call was taken
bool emulate(cublasHandle_t handle, cublasOperation_t ta, cublasOperation_t tb, long long m, long long n, long long k, const void *alpha,
             const void *A, long long lda, const void *B, long long ldb, const void *beta, void *C, long long ldc, int dtype) {
    const Config &cfg = config();
    if (m <= 0 || n <= 0 || k <= 0) return false;
    if ((double)m * (double)n * (double)k < cfg.min_mnk) return false;
    const bool cplx = dtype == GEMMUL8_C32 || dtype == GEMMUL8_C64;
    if (k > (cplx ? (1ll << 16) : (1ll << 17))) return false;
    static auto get_mode = next_symbol<cublasStatus_t (*)(cublasHandle_t, cublasPointerMode_t *)>("cublasGetPointerMode_v2");
    static auto get_stream = next_symbol<cublasStatus_t (*)(cublasHandle_t, cudaStream_t *)>("cublasGetStream_v2");
    cublasPointerMode_t mode;
    if (get_mode(handle, &mode) != CUBLAS_STATUS_SUCCESS) return false;
    cudaStream_t st = nullptr;
    if (get_stream(handle, &st) != CUBLAS_STATUS_SUCCESS) return false;

    gemmul8_b200_args a{};
    a.op_A = op_tag(ta); a.op_B = op_tag(tb);
    a.m = (size_t)m; a.n = (size_t)n; a.k = (size_t)k;
    a.alpha = alpha; a.A = A; a.lda = (size_t)lda; a.B = B; a.ldb = (size_t)ldb; a.beta = beta; a.C = C; a.ldc = (size_t)ldc;
    a.num_moduli   = (dtype == GEMMUL8_F64 || dtype == GEMMUL8_C64) ? cfg.moduli_d : cfg.moduli_s;
    a.fastmode     = cfg.fastmode;
    a.compute_type = cplx ? cfg.complex_type : GEMMUL8_REAL_DEFAULT;
    a.dtype_A = a.dtype_B = a.dtype_C = dtype;
    a.stream = st;
    if (mode == CUBLAS_POINTER_MODE_DEVICE) a.flags |= GEMMUL8_FLAG_DEVICE_SCALARS;
    size_t need = gemmul8_b200_worksize(a.m, a.n, a.k, a.num_moduli, a.compute_type);

    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return false; }
    Workspace *wp = workspace(dev);
    if (!wp) return false;
    Workspace &w = *wp;
    std::lock_guard<std::mutex> lock(w.mu);
    // The full workspace (N (m + n) k bytes and more) may not fit beside the application's matrices, or may exceed the
    // configured cap: real types then take the low-memory call with the largest blocks that do fit (same bits of C).
    size_t block_rows = 0, block_cols = 0, budget = cfg.max_work;
    if (w.bytes < need) {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            const size_t avail = (free_b + w.bytes) / 10 * 9;   // what a re-allocation could get
            if (!budget || avail < budget) budget = avail;
        }
    }
    if (budget && need > budget) {
        if (cplx || gemmul8_b200_plan_blocks(a.m, a.n, a.k, a.num_moduli, budget, &block_rows, &block_cols, &need) != GEMMUL8_OK)
            return false;                                  // nothing fits: let cuBLAS do it
    }
    if (!w.last && cudaEventCreateWithFlags(&w.last, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); w.last = nullptr; return false; }
    if (w.bytes < need) {
        // grow: the previous call on this device (the only user of the old buffer) must have finished with it
        if (w.ptr) {
            if (w.recorded && cudaEventSynchronize(w.last) != cudaSuccess) { cudaGetLastError(); return false; }
            cudaFree(w.ptr); w.ptr = nullptr; w.bytes = 0;
        }
        if (cudaMalloc(&w.ptr, need) != cudaSuccess) { cudaGetLastError(); w.ptr = nullptr; return false; }   // no room: let cuBLAS do it
        w.bytes = need;
    } else if (w.recorded) {
        if (cudaStreamWaitEvent(st, w.last, 0) != cudaSuccess) { cudaGetLastError(); return false; }   // another stream may still be using the buffer
    }
    a.work = w.ptr;
    const int rc = block_rows ? gemmul8_b200_gemm_blocked(&a, block_rows, block_cols) : gemmul8_b200_gemm(&a);
    // whatever was enqueued (a refused call may have enqueued part of its work) is ordered before the next user
    if (cudaEventRecord(w.last, st) == cudaSuccess) w.recorded = true;
    else { cudaGetLastError(); cudaStreamSynchronize(st); w.recorded = false; }
    if (rc != GEMMUL8_OK) {
        if (cfg.verbose) fprintf(stderr, "gemmul8_b200_blas: emulation refused (%s), falling back to cuBLAS\n", gemmul8_b200_last_error());
        return false;
    }
    g_intercepted.fetch_add(1);
    if (cfg.verbose)
        fprintf(stderr, "gemmul8_b200_blas: %lld x %lld x %lld dtype %d -> %u moduli, %s mode%s\n", m, n, k, dtype, a.num_moduli,
                a.fastmode ? "fast" : "accurate", block_rows ? ", low-memory blocks" : "");
    return true;
}

}  // namespace

extern "C" {

unsigned long long gemmul8_b200_blas_intercepted(void) { return g_intercepted.load(); }

cublasStatus_t cublasDgemm_v2(cublasHandle_t h, cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k, const double *alpha,
                              const double *A, int lda, const double *B, int ldb, const double *beta, double *C, int ldc) {
    if (emulate(h, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, GEMMUL8_F64)) return CUBLAS_STATUS_SUCCESS;
    static auto real = next_symbol<decltype(&cublasDgemm_v2)>("cublasDgemm_v2");
    return real(h, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
}
cublasStatus_t cublasSgemm_v2(cublasHandle_t h, cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k, const float *alpha,
                              const float *A, int lda, const float *B, int ldb, const float *beta, float *C, int ldc) {
    if (emulate(h, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, GEMMUL8_F32)) return CUBLAS_STATUS_SUCCESS;
    static auto real = next_symbol<decltype(&cublasSgemm_v2)>("cublasSgemm_v2");
    return real(h, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
}
cublasStatus_t cublasZgemm_v2(cublasHandle_t h, cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k, const cuDoubleComplex *alpha,
                              const cuDoubleComplex *A, int lda, const cuDoubleComplex *B, int ldb, const cuDoubleComplex *beta,
                              cuDoubleComplex *C, int ldc) {
    if (emulate(h, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, GEMMUL8_C64)) return CUBLAS_STATUS_SUCCESS;
    static auto real = next_symbol<decltype(&cublasZgemm_v2)>("cublasZgemm_v2");
    return real(h, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
}
cublasStatus_t cublasCgemm_v2(cublasHandle_t h, cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k, const cuComplex *alpha,
                              const cuComplex *A, int lda, const cuComplex *B, int ldb, const cuComplex *beta, cuComplex *C, int ldc) {
    if (emulate(h, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, GEMMUL8_C32)) return CUBLAS_STATUS_SUCCESS;
    static auto real = next_symbol<decltype(&cublasCgemm_v2)>("cublasCgemm_v2");
    return real(h, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
}
cublasStatus_t cublasGemmEx(cublasHandle_t h, cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k, const void *alpha, const void *A,
                            cudaDataType At, int lda, const void *B, cudaDataType Bt, int ldb, const void *beta, void *C, cudaDataType Ct,
                            int ldc, cublasComputeType_t ct, cublasGemmAlgo_t algo) {
    int dtype = -1;
    if (At == Bt && Bt == Ct) {
        if (At == CUDA_R_64F && ct == CUBLAS_COMPUTE_64F) dtype = GEMMUL8_F64;
        else if (At == CUDA_R_32F && ct == CUBLAS_COMPUTE_32F) dtype = GEMMUL8_F32;
        else if (At == CUDA_C_64F && ct == CUBLAS_COMPUTE_64F) dtype = GEMMUL8_C64;
        else if (At == CUDA_C_32F && ct == CUBLAS_COMPUTE_32F) dtype = GEMMUL8_C32;
    }
    if (dtype >= 0 && emulate(h, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, dtype)) return CUBLAS_STATUS_SUCCESS;
    using gemm_ex_t = cublasStatus_t (*)(cublasHandle_t, cublasOperation_t, cublasOperation_t, int, int, int, const void *, const void *,
                                         cudaDataType, int, const void *, cudaDataType, int, const void *, void *, cudaDataType, int,
                                         cublasComputeType_t, cublasGemmAlgo_t);
    static auto real = next_symbol<gemm_ex_t>("cublasGemmEx");
    return real(h, ta, tb, m, n, k, alpha, A, At, lda, B, Bt, ldb, beta, C, Ct, ldc, ct, algo);
}

}  // extern "C"
