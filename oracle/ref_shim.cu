// ref_shim.cu -- a C-callable door into the UNMODIFIED reference library, for parity tests and the
// `bench.py --impl reference` arm.  TEST INFRASTRUCTURE ONLY (see oracle/oracle.c header).
//
// This file is ours; it is compiled together with the reference's own translation unit
// (/root/reference/GEMMul8/src/gemmul8.cu, where it lies) by oracle/Makefile into
// oracle/_ref/libgemmul8_ref.so.  It only forwards to gemmul8::workSize / gemmul8::gemm<TA,TB,TC>
// (GEMMul8/include/gemmul8.hpp:18-287) so tests can drive the reference through ctypes and then
// read its workspace (A8i | B8i | C8u | C32i | sftA | sftB, GEMMul8/src/gemmul8.cu:229-234).
#include "gemmul8.hpp"

#include <cstdio>

namespace {
cublasHandle_t handle() {
    static cublasHandle_t h = nullptr;
    if (!h) cublasCreate(&h);
    return h;
}
template <typename TA, typename TB, typename TC>
int call(int opA, int opB, size_t m, size_t n, size_t k, const void *alpha, const void *A, size_t lda, const void *B, size_t ldb,
         const void *beta, void *C, size_t ldc, unsigned N, int fast, void *work, int ct, double *timers) {
    std::vector<double> t = gemmul8::gemm<TA, TB, TC>(handle(), (cublasOperation_t)opA, (cublasOperation_t)opB, m, n, k,
                                                      static_cast<const TC *>(alpha), static_cast<const TA *>(A), lda,
                                                      static_cast<const TB *>(B), ldb, static_cast<const TC *>(beta),
                                                      static_cast<TC *>(C), ldc, N, fast != 0, work, (gemmul8::computeType_t)ct);
    if (timers) for (int i = 0; i < 4; ++i) timers[i] = t[i];
    return cudaGetLastError() == cudaSuccess ? 0 : 3;
}
}  // namespace

extern "C" {

size_t gemmul8_ref_worksize(size_t m, size_t n, size_t k, unsigned N, int ct) {
    return gemmul8::workSize(m, n, k, N, (gemmul8::computeType_t)ct);
}

// dtype tags: 0 f32, 1 f64, 2 c32, 3 c64
int gemmul8_ref_gemm(int dtA, int dtB, int dtC, int opA, int opB, size_t m, size_t n, size_t k, const void *alpha, const void *A,
                     size_t lda, const void *B, size_t ldb, const void *beta, void *C, size_t ldc, unsigned N, int fast, void *work,
                     int ct, double *timers) {
#define OZ_CASE(a, b, c, TA, TB, TC) \
    if (dtA == a && dtB == b && dtC == c) return call<TA, TB, TC>(opA, opB, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, N, fast, work, ct, timers);
    OZ_CASE(1, 1, 1, double, double, double)
    OZ_CASE(0, 0, 0, float, float, float)
    OZ_CASE(1, 0, 1, double, float, double)
    OZ_CASE(0, 1, 1, float, double, double)
    OZ_CASE(1, 0, 0, double, float, float)
    OZ_CASE(0, 1, 0, float, double, float)
    OZ_CASE(2, 2, 2, cuFloatComplex, cuFloatComplex, cuFloatComplex)
    OZ_CASE(3, 3, 3, cuDoubleComplex, cuDoubleComplex, cuDoubleComplex)
    OZ_CASE(3, 2, 3, cuDoubleComplex, cuFloatComplex, cuDoubleComplex)
    OZ_CASE(2, 3, 3, cuFloatComplex, cuDoubleComplex, cuDoubleComplex)
    OZ_CASE(3, 2, 2, cuDoubleComplex, cuFloatComplex, cuFloatComplex)
    OZ_CASE(2, 3, 2, cuFloatComplex, cuDoubleComplex, cuFloatComplex)
#undef OZ_CASE
    fprintf(stderr, "gemmul8_ref_gemm: the reference exports no specialisation for dtypes (%d,%d,%d)\n", dtA, dtB, dtC);
    return 2;
}

}  // extern "C"
