// gemmul8.hpp -- C++ drop-in for the public API of ptrkgtsch/mixed-GEMMul8 on B200.
//
// Same namespace, enum order, function names, argument order and defaults as the reference header
// (GEMMul8/include/gemmul8.hpp:7-12 computeType_t, :18-22 workSize, :29-47 gemm<TA,TB,TC>, :49-287 its
// 12 explicit specialisations), so existing callers -- the reference's own drivers under
// GEMMul8/testing/, or user code that links -lgemmul8 -- recompile against this header and link
// libgemmul8_b200.so instead, unchanged.  The specialisations are compiled into that library with
// the SAME mangled names as the reference's libgemmul8.a (link-level drop-in); they forward to the
// C ABI of include/gemmul8_b200.h.
//
// Behaviour kept from the reference: column-major device matrices, HOST alpha/beta, `work` of at
// least workSize() bytes, the returned vector holds 4 phase times in nanoseconds (the reference's
// header says seconds, its code stores ns: GEMMul8/src/gemmul8.cu:17), the call has completed when
// it returns, an unsupported computeType prints a message and returns zeros with C untouched
// (GEMMul8/src/gemmul8.cu:174-177).  Differences: kernels run on the handle's stream (the reference
// ignores it and uses the legacy default stream), and the call is thread-safe.
// Define GEMMUL8_B200_ASYNC before including to get a fully asynchronous call (returns zeros).
#pragma once
#include <vector>
#include <cstddef>
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cuComplex.h>

// the reference's portability names (GEMMul8/src/gpu_arch.hpp:12-38), CUDA meanings only
#ifndef gpublasHandle_t
#define gpublasHandle_t cublasHandle_t
#define gpublasOperation_t cublasOperation_t
#define GPUBLAS_OP_N CUBLAS_OP_N
#define GPUBLAS_OP_T CUBLAS_OP_T
#define GPUBLAS_OP_C CUBLAS_OP_C
#define gpuDoubleComplex cuDoubleComplex
#define gpuFloatComplex cuFloatComplex
#define gpuDeviceSynchronize cudaDeviceSynchronize
#define make_gpuFloatComplex make_cuFloatComplex
#define make_gpuDoubleComplex make_cuDoubleComplex
#define gpuCreal cuCreal
#define gpuCrealf cuCrealf
#define gpuCimag cuCimag
#define gpuCimagf cuCimagf
#endif

namespace gemmul8 {

typedef enum {
    REAL_DEFAULT,
    COMPLEX_BIG_MATRIX_ENCODE,
    COMPLEX_CLASSIC_MULT,
    COMPLEX_KARATSUBA_MULT
} computeType_t;

// bytes of device scratch a gemm call needs (same formula as the reference, so callers' buffers fit)
size_t workSize(const size_t m, const size_t n, const size_t k, const unsigned num_moduli,
                const computeType_t computeType = REAL_DEFAULT);

// C = alpha * op(A) * op(B) + beta * C by Ozaki scheme II on the int8 tensor cores.
// 2 <= num_moduli <= 20, k <= 2^17; fastmode = true (vector-norm bound) or false (int8 bound product).
template <typename TA, typename TB = TA, typename TC = TA>
std::vector<double> gemm(gpublasHandle_t handle, const gpublasOperation_t op_A, const gpublasOperation_t op_B,
                         const size_t m, const size_t n, const size_t k, const TC *alpha, const TA *const A,
                         const size_t lda, const TB *const B, const size_t ldb, const TC *beta, TC *const C,
                         const size_t ldc, const unsigned num_moduli, const bool fastmode, void *const work,
                         const computeType_t computeType = REAL_DEFAULT);

#define GEMMUL8_B200_DECLARE(TA, TB, TC)                                                                             \
    template <>                                                                                                      \
    std::vector<double> gemm<TA, TB, TC>(gpublasHandle_t handle, const gpublasOperation_t op_A,                      \
                                         const gpublasOperation_t op_B, const size_t m, const size_t n,              \
                                         const size_t k, const TC *alpha, const TA *const A, const size_t lda,       \
                                         const TB *const B, const size_t ldb, const TC *beta, TC *const C,           \
                                         const size_t ldc, const unsigned num_moduli, const bool fastmode,           \
                                         void *const work, const computeType_t computeType);
GEMMUL8_B200_DECLARE(double, double, double)
GEMMUL8_B200_DECLARE(float, float, float)
GEMMUL8_B200_DECLARE(double, float, double)
GEMMUL8_B200_DECLARE(float, double, double)
GEMMUL8_B200_DECLARE(double, float, float)
GEMMUL8_B200_DECLARE(float, double, float)
GEMMUL8_B200_DECLARE(gpuFloatComplex, gpuFloatComplex, gpuFloatComplex)
GEMMUL8_B200_DECLARE(gpuDoubleComplex, gpuDoubleComplex, gpuDoubleComplex)
GEMMUL8_B200_DECLARE(gpuDoubleComplex, gpuFloatComplex, gpuDoubleComplex)
GEMMUL8_B200_DECLARE(gpuFloatComplex, gpuDoubleComplex, gpuDoubleComplex)
GEMMUL8_B200_DECLARE(gpuDoubleComplex, gpuFloatComplex, gpuFloatComplex)
GEMMUL8_B200_DECLARE(gpuFloatComplex, gpuDoubleComplex, gpuFloatComplex)
#undef GEMMUL8_B200_DECLARE

}  // namespace gemmul8
