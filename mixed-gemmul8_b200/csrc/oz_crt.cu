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
#include "oz_crt.cuh"

namespace oz {
namespace {

// N is a compile-time constant: the modulus loop unrolls completely, the CRT weights become
// constant-bank operands of the DFMAs and all N residue loads are issued before the first use.
template <typename T, bool SPLIT, int N>
__global__ void __launch_bounds__(256) crt_kernel(size_t m, size_t n, const uint8_t *__restrict__ C8u, size_t ldc8u,
                                                  size_t sizeC, T *__restrict__ C, size_t ldc,
                                                  const int16_t *__restrict__ sftA, const int16_t *__restrict__ sftB,
                                                  int mode, T alpha, T beta, const T *__restrict__ alpha_dev,
                                                  const T *__restrict__ beta_dev) {
    if (alpha_dev != nullptr) {   // CUBLAS_POINTER_MODE_DEVICE-style scalars: read here, never on the host
        alpha = *alpha_dev; beta = *beta_dev;
        mode  = alpha_beta_mode(alpha, beta);
    }
    const size_t row0 = ((size_t)blockIdx.x * 64 + threadIdx.x) * 4;
    const size_t col  = (size_t)blockIdx.y * 4 + threadIdx.y;
    if (row0 >= m || col >= n) return;
    uint32_t res[N];
    const uint8_t *__restrict__ src = C8u + col * ldc8u + row0;
#pragma unroll
    for (int j = 0; j < N; ++j) res[j] = *reinterpret_cast<const uint32_t *>(src + (size_t)j * sizeC);

    double s1[4] = {0.0, 0.0, 0.0, 0.0}, s2[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int j = 0; j < N; ++j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) crt_step<SPLIT>(N, j, __byte_perm(res[j], 0, 0x4440 + e), s1[e], s2[e]);
    }
    const int sb = sftB[col];
    T *cptr = C + col * ldc + row0;
    if (row0 + 3 < m) {
        int sa[4];
        if ((reinterpret_cast<uintptr_t>(sftA + row0) & 7) == 0) {
            const short4 s4 = *reinterpret_cast<const short4 *>(sftA + row0);
            sa[0] = s4.x; sa[1] = s4.y; sa[2] = s4.z; sa[3] = s4.w;
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) sa[e] = sftA[row0 + e];
        }
        T out[4];
        int ex[4];
        bool normal = true;
#pragma unroll
        for (int e = 0; e < 4; ++e) { ex[e] = sa[e] + sb; normal &= (unsigned)(ex[e] + 1022) <= 2045u; }
        if (mode == AB_10 && normal) {
            // the common case (alpha = 1, beta = 0, every 2^e a normal number) without the per-element mode dispatch
            // and scalbn fallbacks: one exact multiply by a constructed power of two
#pragma unroll
            for (int e = 0; e < 4; ++e)
                out[e] = cast_out<T>(crt_finish<SPLIT>(N, s1[e], s2[e]) * __hiloint2double((1023 + ex[e]) << 20, 0));
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                out[e] = combine<T>(mode, alpha, beta, cast_out<T>(scale_pow2(crt_finish<SPLIT>(N, s1[e], s2[e]), ex[e])), cptr + e);
        }
        if ((reinterpret_cast<uintptr_t>(cptr) & (4 * sizeof(T) - 1)) == 0) {
            if constexpr (sizeof(T) == 8) {
                reinterpret_cast<double2 *>(cptr)[0] = make_double2(out[0], out[1]);
                reinterpret_cast<double2 *>(cptr)[1] = make_double2(out[2], out[3]);
            } else {
                *reinterpret_cast<float4 *>(cptr) = make_float4(out[0], out[1], out[2], out[3]);
            }
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) cptr[e] = out[e];
        }
    } else {
        for (int e = 0; e < 4 && row0 + e < m; ++e)
            cptr[e] = combine<T>(mode, alpha, beta, cast_out<T>(scale_pow2(crt_finish<SPLIT>(N, s1[e], s2[e]), (int)sftA[row0 + e] + sb)), cptr + e);
    }
}

template <typename T, bool SPLIT, int N>
void launch_one(dim3 grid, dim3 block, cudaStream_t st, size_t m, size_t n, const uint8_t *C8u, size_t ldc8u, size_t sizeC,
                void *C, size_t ldc, const int16_t *sftA, const int16_t *sftB, int mode, T alpha, T beta, const T *alpha_dev,
                const T *beta_dev) {
    crt_kernel<T, SPLIT, N><<<grid, block, 0, st>>>(m, n, C8u, ldc8u, sizeC, static_cast<T *>(C), ldc, sftA, sftB, mode, alpha, beta,
                                                    alpha_dev, beta_dev);
}

template <typename T>
cudaError_t run_crt(bool split, unsigned N, size_t m, size_t n, const uint8_t *C8u, size_t ldc8u, size_t sizeC, void *C,
                    size_t ldc, const int16_t *sftA, const int16_t *sftB, const void *alpha_p, const void *beta_p,
                    bool device_scalars, cudaStream_t st) {
    const T *alpha_dev = device_scalars ? static_cast<const T *>(alpha_p) : nullptr;
    const T *beta_dev  = device_scalars ? static_cast<const T *>(beta_p) : nullptr;
    const T alpha = device_scalars ? T(1) : *static_cast<const T *>(alpha_p), beta = device_scalars ? T(0) : *static_cast<const T *>(beta_p);
    const int mode = alpha_beta_mode(alpha, beta);
    dim3 block(64, 4), grid((unsigned)(((m + 3) / 4 + 63) / 64), (unsigned)((n + 3) / 4));
#define OZ_CRT_CASE(NN)                                                                                                          \
    case NN:                                                                                                                     \
        if constexpr (NN >= 8 && sizeof(T) == 8) {                                                                               \
            if (split) { launch_one<T, true, NN>(grid, block, st, m, n, C8u, ldc8u, sizeC, C, ldc, sftA, sftB, mode, alpha, beta, alpha_dev, beta_dev); break; } \
        }                                                                                                                        \
        launch_one<T, false, NN>(grid, block, st, m, n, C8u, ldc8u, sizeC, C, ldc, sftA, sftB, mode, alpha, beta, alpha_dev, beta_dev); \
        break;
    switch (N) {
        OZ_CRT_CASE(2) OZ_CRT_CASE(3) OZ_CRT_CASE(4) OZ_CRT_CASE(5) OZ_CRT_CASE(6) OZ_CRT_CASE(7) OZ_CRT_CASE(8)
        OZ_CRT_CASE(9) OZ_CRT_CASE(10) OZ_CRT_CASE(11) OZ_CRT_CASE(12) OZ_CRT_CASE(13) OZ_CRT_CASE(14) OZ_CRT_CASE(15)
        OZ_CRT_CASE(16) OZ_CRT_CASE(17) OZ_CRT_CASE(18) OZ_CRT_CASE(19) OZ_CRT_CASE(20)
        default: return cudaErrorInvalidValue;
    }
#undef OZ_CRT_CASE
    count_launch();
    return cudaGetLastError();
}

}  // namespace

namespace {
// k == 0: the product is empty, C = beta * C (BLAS: C is not read when beta == 0)
template <typename T>
__global__ void scale_c_kernel(size_t m, size_t n, T *__restrict__ C, size_t ldc, T beta_re, T beta_im, bool cplx,
                               const T *__restrict__ beta_dev) {
    if (beta_dev != nullptr) { beta_re = beta_dev[0]; beta_im = cplx ? beta_dev[1] : T(0); }
    const size_t w = cplx ? 2 * m : m;                 // scalars per column
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, col = blockIdx.y;
    if (i >= w) return;
    T *c = C + col * (cplx ? 2 * ldc : ldc);
    if (beta_re == T(0) && beta_im == T(0)) { c[i] = T(0); return; }
    if (!cplx) { c[i] = beta_re * c[i]; return; }
    if (i & 1) return;                                 // even lanes own (re, im)
    const T re = c[i], im = c[i + 1];
    c[i]     = beta_re * re - beta_im * im;
    c[i + 1] = beta_re * im + beta_im * re;
}
}  // namespace

cudaError_t launch_scale_c(int dtype_C, size_t m, size_t n, void *C, size_t ldc, const void *beta_p, bool device_scalars,
                           cudaStream_t st) {
    if (m == 0 || n == 0) return cudaSuccess;
    if (n > 65535) return cudaErrorInvalidValue;
    const bool cplx = dtype_C == DT_C32 || dtype_C == DT_C64;
    dim3 grid((unsigned)(((cplx ? 2 * m : m) + 255) / 256), (unsigned)n);
    if (dtype_C == DT_F64 || dtype_C == DT_C64) {
        const double *b = static_cast<const double *>(beta_p);
        if (device_scalars) scale_c_kernel<double><<<grid, 256, 0, st>>>(m, n, static_cast<double *>(C), ldc, 0.0, 0.0, cplx, b);
        else scale_c_kernel<double><<<grid, 256, 0, st>>>(m, n, static_cast<double *>(C), ldc, b[0], cplx ? b[1] : 0.0, cplx, nullptr);
    } else {
        const float *b = static_cast<const float *>(beta_p);
        if (device_scalars) scale_c_kernel<float><<<grid, 256, 0, st>>>(m, n, static_cast<float *>(C), ldc, 0.f, 0.f, cplx, b);
        else scale_c_kernel<float><<<grid, 256, 0, st>>>(m, n, static_cast<float *>(C), ldc, b[0], cplx ? b[1] : 0.f, cplx, nullptr);
    }
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_crt(int dtype_C, bool split_weights, unsigned num_moduli, size_t m, size_t n, const uint8_t *C8u,
                       size_t ldc8u, size_t sizeC, void *C, size_t ldc, const int16_t *sftA, const int16_t *sftB,
                       const void *alpha_p, const void *beta_p, bool device_scalars, cudaStream_t st) {
    if (m == 0 || n == 0) return cudaSuccess;
    if ((n + 3) / 4 > 65535) return cudaErrorInvalidValue;
    switch (dtype_C) {
        case DT_F64: return run_crt<double>(split_weights, num_moduli, m, n, C8u, ldc8u, sizeC, C, ldc, sftA, sftB, alpha_p, beta_p, device_scalars, st);
        case DT_F32: return run_crt<float>(false, num_moduli, m, n, C8u, ldc8u, sizeC, C, ldc, sftA, sftB, alpha_p, beta_p, device_scalars, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace oz
