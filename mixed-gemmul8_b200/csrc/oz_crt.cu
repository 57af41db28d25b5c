// CRT accumulation, reduction mod M, inverse power-of-two scaling and alpha/beta.
//
// Arithmetic follows the reference so that C is bit-identical on its tested path:
//   single weights (N <= 7, or fp32 output)  GEMMul8/src/inverse_scaling.hpp:35-62
//   split  weights (N >= 8, fp64 output)     GEMMul8/src/inverse_scaling.hpp:140-172
//   which one: is_numM_1                     GEMMul8/src/gemmul8.cu:201-202, :486
// alpha/beta are applied BLAS-correctly, C = alpha*c + beta*C, with the reference's FMA shapes
// where the reference is itself correct ((1,0), (1,1), (a,1), (a,b): inverse_scaling.hpp:268-820).
// The reference's (1,b) kernels compute beta*c + C (inverse_scaling.hpp:417,682) and its split
// (a,1) kernel alpha*C + c (:736); those are defects we do not reproduce.  C is not read when beta == 0.
//
// Layout: one thread = 4 consecutive rows of one column.  A warp reads 128 contiguous residue
// bytes per modulus and writes 1 KiB (fp64) of C; all N residue loads are issued before use.
#include "oz_common.cuh"

namespace oz {
namespace {

template <typename T> __device__ __forceinline__ T cast_out(double v);
template <> __device__ __forceinline__ double cast_out<double>(double v) { return v; }
template <> __device__ __forceinline__ float cast_out<float>(double v) { return __double2float_rn(v); }

__device__ __forceinline__ double fma_t(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float fma_t(float a, float b, float c) { return __fmaf_rn(a, b, c); }

enum AlphaBeta : int { AB_10 = 0, AB_11, AB_1B, AB_A0, AB_A1, AB_AB };

template <typename T>
__device__ __forceinline__ T combine(int mode, T alpha, T beta, T c, const T *cptr) {
    switch (mode) {
        case AB_10: return c;
        case AB_11: return c + *cptr;
        case AB_1B: return fma_t(beta, *cptr, c);
        case AB_A0: return alpha * c;
        case AB_A1: return fma_t(alpha, c, *cptr);
        default:    return fma_t(beta, *cptr, alpha * c);
    }
}

template <typename T, bool SPLIT>
__global__ void __launch_bounds__(256) crt_kernel(unsigned num_moduli, size_t m, size_t n,
                                                  const uint8_t *__restrict__ C8u, size_t ldc8u, size_t sizeC,
                                                  T *__restrict__ C, size_t ldc, const int16_t *__restrict__ sftA,
                                                  const int16_t *__restrict__ sftB, int mode, T alpha, T beta) {
    const size_t row0 = ((size_t)blockIdx.x * 64 + threadIdx.x) * 4;
    const size_t col  = (size_t)blockIdx.y * 4 + threadIdx.y;
    if (row0 >= m || col >= n) return;
    const unsigned ti = num_moduli - 2;

    uint32_t res[kMaxModuli];
    const uint8_t *__restrict__ src = C8u + col * ldc8u + row0;
#pragma unroll
    for (int j = 0; j < kMaxModuli; ++j)
        if (j < (int)num_moduli) res[j] = *reinterpret_cast<const uint32_t *>(src + (size_t)j * sizeC);

    double s1[4] = {0.0, 0.0, 0.0, 0.0}, s2[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int j = 0; j < kMaxModuli; ++j) {
        if (j < (int)num_moduli) {
            double w1, w2 = 0.0;
            if constexpr (SPLIT) { w1 = dev_tab::OZ_W2_HI[num_moduli - 8][j]; w2 = dev_tab::OZ_W2_LO[num_moduli - 8][j]; }
            else { w1 = dev_tab::OZ_W1[ti][j]; }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const double r = __uint2double_rn((res[j] >> (8 * e)) & 0xffu);
                s1[e] = fma(w1, r, s1[e]);
                if constexpr (SPLIT) s2[e] = fma(w2, r, s2[e]);
            }
        }
    }
    const double invM = dev_tab::OZ_INV_M[ti], M1 = dev_tab::OZ_M_HI[ti], M2 = dev_tab::OZ_M_LO[ti];
    const int sb = sftB[col];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const size_t row = row0 + e;
        if (row >= m) break;
        double c;
        if constexpr (SPLIT) {
            const double quot = -rint(fma(s1[e], invM, s2[e] * invM));
            const double t1   = fma(quot, M1, s1[e]) + s2[e];
            c                 = fma(quot, M2, t1);
        } else {
            const double quot = -rint(s1[e] * invM);
            c                 = fma(quot, M1, s1[e]);
        }
        c = scalbn(c, (int)sftA[row] + sb);
        T *cptr = C + col * ldc + row;
        *cptr   = combine<T>(mode, alpha, beta, cast_out<T>(c), cptr);
    }
}

template <typename T>
cudaError_t run_crt(bool split, unsigned N, size_t m, size_t n, const uint8_t *C8u, size_t ldc8u, size_t sizeC, void *C,
                    size_t ldc, const int16_t *sftA, const int16_t *sftB, const void *alpha_host, const void *beta_host,
                    cudaStream_t st) {
    const T alpha = *static_cast<const T *>(alpha_host), beta = *static_cast<const T *>(beta_host);
    int mode;
    if (alpha == T(1)) mode = (beta == T(0)) ? AB_10 : (beta == T(1)) ? AB_11 : AB_1B;
    else               mode = (beta == T(0)) ? AB_A0 : (beta == T(1)) ? AB_A1 : AB_AB;
    dim3 block(64, 4), grid((unsigned)(((m + 3) / 4 + 63) / 64), (unsigned)((n + 3) / 4));
    if (split) crt_kernel<T, true><<<grid, block, 0, st>>>(N, m, n, C8u, ldc8u, sizeC, static_cast<T *>(C), ldc, sftA, sftB, mode, alpha, beta);
    else       crt_kernel<T, false><<<grid, block, 0, st>>>(N, m, n, C8u, ldc8u, sizeC, static_cast<T *>(C), ldc, sftA, sftB, mode, alpha, beta);
    count_launch();
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_crt(int dtype_C, bool split_weights, unsigned num_moduli, size_t m, size_t n, const uint8_t *C8u,
                       size_t ldc8u, size_t sizeC, void *C, size_t ldc, const int16_t *sftA, const int16_t *sftB,
                       const void *alpha_host, const void *beta_host, cudaStream_t st) {
    if (m == 0 || n == 0) return cudaSuccess;
    if ((n + 3) / 4 > 65535) return cudaErrorInvalidValue;
    switch (dtype_C) {
        case DT_F64: return run_crt<double>(split_weights, num_moduli, m, n, C8u, ldc8u, sizeC, C, ldc, sftA, sftB, alpha_host, beta_host, st);
        case DT_F32: return run_crt<float>(false, num_moduli, m, n, C8u, ldc8u, sizeC, C, ldc, sftA, sftB, alpha_host, beta_host, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace oz
