// CRT accumulation / reduction mod M / alpha-beta combine, shared by the stand-alone CRT kernel
// (oz_crt.cu) and the fused epilogue of the all-moduli GEMM (oz_gemm.cu).
//
// Arithmetic follows the reference so that C is bit-identical on its tested path:
//   single weights (N <= 7, or fp32 output)  GEMMul8/src/inverse_scaling.hpp:35-62
//   split  weights (N >= 8, fp64 output)     GEMMul8/src/inverse_scaling.hpp:140-172
#pragma once
#include "oz_common.cuh"

#ifndef __CUDACC__
#define __host__
#define __device__
#endif

namespace oz {

enum AlphaBeta : int { AB_10 = 0, AB_11, AB_1B, AB_A0, AB_A1, AB_AB };

template <typename T> __host__ __device__ inline int alpha_beta_mode(T alpha, T beta) {
    if (alpha == T(1)) return (beta == T(0)) ? AB_10 : (beta == T(1)) ? AB_11 : AB_1B;
    return (beta == T(0)) ? AB_A0 : (beta == T(1)) ? AB_A1 : AB_AB;
}

#ifdef __CUDACC__
template <typename T> __device__ __forceinline__ T cast_out(double v);
template <> __device__ __forceinline__ double cast_out<double>(double v) { return v; }
template <> __device__ __forceinline__ float cast_out<float>(double v) { return __double2float_rn(v); }

__device__ __forceinline__ double fma_t(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float fma_t(float a, float b, float c) { return __fmaf_rn(a, b, c); }

// BLAS-correct C = alpha*c + beta*C with the reference's FMA shapes where the reference is itself
// correct ((1,0), (1,1), (a,1), (a,b): inverse_scaling.hpp:268-820); C is not read when beta == 0.
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

// exact uint8 -> double without the conversion pipe: bits of 2^52 + r, minus 2^52
__device__ __forceinline__ double byte_to_double(uint32_t r) {
    return __hiloint2double(0x43300000, (int)r) - 4503599627370496.0;
}

// c * 2^e: one exact multiplication when 2^e is a normal number (scalbn otherwise)
__device__ __forceinline__ double scale_pow2(double c, int e) {
    if (e >= -1022 && e <= 1023) return c * __hiloint2double((1023 + e) << 20, 0);
    return scalbn(c, e);
}

// one CRT step: s1 += w1 * r (and s2 += w2 * r with split weights)
template <bool SPLIT>
__device__ __forceinline__ void crt_step(unsigned num_moduli, int j, uint32_t r, double &s1, double &s2) {
    const double rd = byte_to_double(r);
    if constexpr (SPLIT) {
        s1 = fma(dev_tab::OZ_W2_HI[num_moduli - 8][j], rd, s1);
        s2 = fma(dev_tab::OZ_W2_LO[num_moduli - 8][j], rd, s2);
    } else {
        s1 = fma(dev_tab::OZ_W1[num_moduli - 2][j], rd, s1);
    }
}

// reduction mod M to the symmetric representative (still scaled by 2^-(sftA+sftB))
template <bool SPLIT>
__device__ __forceinline__ double crt_finish(unsigned num_moduli, double s1, double s2) {
    const unsigned ti = num_moduli - 2;
    const double invM = dev_tab::OZ_INV_M[ti], M1 = dev_tab::OZ_M_HI[ti];
    if constexpr (SPLIT) {
        const double M2   = dev_tab::OZ_M_LO[ti];
        const double quot = -rint(fma(s1, invM, s2 * invM));
        const double t1   = fma(quot, M1, s1) + s2;
        return fma(quot, M2, t1);
    } else {
        const double quot = -rint(s1 * invM);
        return fma(quot, M1, s1);
    }
}
#endif

}  // namespace oz
