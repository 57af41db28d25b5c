// Link-level drop-in: definitions of gemmul8::workSize and the 12 gemmul8::gemm<TA,TB,TC>
// specialisations with the reference's mangled names (GEMMul8/src/gemmul8.cu:129-147, :149-431,
// :1054-1316), each a thin forward to the C ABI (include/gemmul8_b200.h).
//
// cuBLAS is only needed for one thing -- asking the caller's handle for its stream -- and the
// product library does not link it: cublasGetStream_v2 is looked up in the running process (a
// caller that owns a handle has cuBLAS loaded).  Without it the legacy default stream is used, as
// the reference does.
#include "../../include/gemmul8.hpp"
#include "../../include/gemmul8_b200.h"

#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

namespace {

cudaStream_t stream_of(cublasHandle_t handle) {
    if (!handle) return nullptr;
    using fn_t = cublasStatus_t (*)(cublasHandle_t, cudaStream_t *);
    static fn_t fn = reinterpret_cast<fn_t>(dlsym(RTLD_DEFAULT, "cublasGetStream_v2"));
    cudaStream_t st = nullptr;
    if (fn && fn(handle, &st) == CUBLAS_STATUS_SUCCESS) return st;
    return nullptr;
}

template <typename T> constexpr int dtype_tag() {
    if (std::is_same<T, float>::value) return GEMMUL8_F32;
    if (std::is_same<T, double>::value) return GEMMUL8_F64;
    if (std::is_same<T, cuFloatComplex>::value) return GEMMUL8_C32;
    return GEMMUL8_C64;
}

template <typename TA, typename TB, typename TC>
std::vector<double> forward(cublasHandle_t handle, cublasOperation_t op_A, cublasOperation_t op_B, size_t m, size_t n,
                            size_t k, const TC *alpha, const TA *A, size_t lda, const TB *B, size_t ldb, const TC *beta,
                            TC *C, size_t ldc, unsigned num_moduli, bool fastmode, void *work,
                            gemmul8::computeType_t computeType, bool async) {
    gemmul8_b200_args a{};
    a.op_A = (int)op_A; a.op_B = (int)op_B;
    a.m = m; a.n = n; a.k = k;
    a.alpha = alpha; a.A = A; a.lda = lda; a.B = B; a.ldb = ldb; a.beta = beta; a.C = C; a.ldc = ldc;
    a.num_moduli = num_moduli; a.fastmode = fastmode ? 1 : 0; a.work = work; a.compute_type = (int)computeType;
    a.dtype_A = dtype_tag<TA>(); a.dtype_B = dtype_tag<TB>(); a.dtype_C = dtype_tag<TC>();
    a.stream = stream_of(handle);
    a.flags  = async ? 0u : (unsigned)GEMMUL8_FLAG_TIMERS;   // timers synchronise: the call is complete on return
    const int rc = gemmul8_b200_gemm(&a);
    std::vector<double> t(4, 0.0);
    if (rc == GEMMUL8_OK) {
        for (int i = 0; i < 4; ++i) t[i] = a.timers_ns[i];
    } else if (rc != GEMMUL8_ERR_COMPUTETYPE) {   // that one already printed the reference's message
        fprintf(stderr, "gemmul8::gemm: %s\n", gemmul8_b200_last_error());
    }
    return t;
}

bool async_default() {
    static const bool v = getenv("GEMMUL8_B200_ASYNC") != nullptr;
    return v;
}

}  // namespace

namespace gemmul8 {

size_t workSize(const size_t m, const size_t n, const size_t k, const unsigned num_moduli, const computeType_t computeType) {
    return gemmul8_b200_worksize(m, n, k, num_moduli, (int)computeType);
}

#define GEMMUL8_B200_DEFINE(TA, TB, TC)                                                                              \
    template <>                                                                                                      \
    std::vector<double> gemm<TA, TB, TC>(gpublasHandle_t handle, const gpublasOperation_t op_A,                      \
                                         const gpublasOperation_t op_B, const size_t m, const size_t n,              \
                                         const size_t k, const TC *alpha, const TA *const A, const size_t lda,       \
                                         const TB *const B, const size_t ldb, const TC *beta, TC *const C,           \
                                         const size_t ldc, const unsigned num_moduli, const bool fastmode,           \
                                         void *const work, const computeType_t computeType) {                        \
        return forward<TA, TB, TC>(handle, op_A, op_B, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, num_moduli,    \
                                   fastmode, work, computeType, async_default());                                    \
    }
GEMMUL8_B200_DEFINE(double, double, double)
GEMMUL8_B200_DEFINE(float, float, float)
GEMMUL8_B200_DEFINE(double, float, double)
GEMMUL8_B200_DEFINE(float, double, double)
GEMMUL8_B200_DEFINE(double, float, float)
GEMMUL8_B200_DEFINE(float, double, float)
GEMMUL8_B200_DEFINE(gpuFloatComplex, gpuFloatComplex, gpuFloatComplex)
GEMMUL8_B200_DEFINE(gpuDoubleComplex, gpuDoubleComplex, gpuDoubleComplex)
GEMMUL8_B200_DEFINE(gpuDoubleComplex, gpuFloatComplex, gpuDoubleComplex)
GEMMUL8_B200_DEFINE(gpuFloatComplex, gpuDoubleComplex, gpuDoubleComplex)
GEMMUL8_B200_DEFINE(gpuDoubleComplex, gpuFloatComplex, gpuFloatComplex)
GEMMUL8_B200_DEFINE(gpuFloatComplex, gpuDoubleComplex, gpuFloatComplex)
#undef GEMMUL8_B200_DEFINE

}  // namespace gemmul8
