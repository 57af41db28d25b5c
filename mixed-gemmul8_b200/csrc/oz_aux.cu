// Measurement helpers for tests and bench.py ONLY (not part of the product library):
//   * the reference's synthetic "phi" matrices, reproduced bit-for-bit on the device
//     (GEMMul8/testing/make_matrix.hpp:7-57: cuRAND XORWOW, curand_init(seed, idx, 0),
//      (U - 0.5) * exp(phi * Z); complex: re then im from the same state);
//   * a double-double "truth" product C1 + C2 ~= A*B used for the relerr columns
//     (the reference uses eval::dd_gpu::simple_gemm, GEMMul8/testing/eval.hpp:265-308).
#include <cuda_runtime.h>
#include <curand_kernel.h>
#include <stdint.h>

namespace {

template <typename T> __device__ __forceinline__ T phi_value(curandState *s, T phi) {
    const T u = static_cast<T>(curand_uniform_double(s));
    const T z = static_cast<T>(curand_normal_double(s));
    return static_cast<T>((u - 0.5) * exp(z * phi));
}

template <typename T, bool CPLX>
__global__ void phi_matrix_kernel(size_t count, T *out, T phi, unsigned long long seed) {
    const size_t idx = threadIdx.x + (size_t)blockIdx.x * blockDim.x;
    if (idx >= count) return;
    curandState st;
    curand_init(seed, idx, 0, &st);
    if (CPLX) {
        const T re = phi_value<T>(&st, phi);
        const T im = phi_value<T>(&st, phi);
        out[2 * idx]     = re;
        out[2 * idx + 1] = im;
    } else {
        out[idx] = phi_value<T>(&st, phi);
    }
}

// error-free transformations
__device__ __forceinline__ void two_sum(double a, double b, double &s, double &e) {
    s = a + b;
    const double bb = s - a;
    e = (a - (s - bb)) + (b - bb);
}
__device__ __forceinline__ void dd_accumulate(double p, double pe, double &hi, double &lo) {
    double s, e;
    two_sum(hi, p, s, e);
    e += lo + pe;
    hi = s + e;
    lo = e - (hi - s);
}

// C1 + C2 = op(A) * op(B) in double-double; column-major, 32x32 tiles, one element per thread.
// Optional row / column subsets (rows[i], cols[j]) evaluate only a sample of C.
__global__ void dd_gemm_kernel(size_t m, size_t n, size_t k, const double *A, size_t lda, int transA, const double *B,
                               size_t ldb, int transB, const int *rows, const int *cols, double *C1, double *C2, size_t ldc) {
    __shared__ double As[32][33], Bs[32][33];
    const size_t ti = blockIdx.x * 32 + threadIdx.x, tj = blockIdx.y * 32 + threadIdx.y;
    // element this thread loads / computes
    const size_t lr = blockIdx.x * 32 + threadIdx.y;  // A tile is loaded with threadIdx.y selecting the row
    const size_t arow = lr < m ? (rows ? (size_t)rows[lr] : lr) : 0;
    const size_t bcol = tj < n ? (cols ? (size_t)cols[tj] : tj) : 0;
    double hi = 0.0, lo = 0.0;
    for (size_t k0 = 0; k0 < k; k0 += 32) {
        const size_t ka = k0 + threadIdx.x, kb = k0 + threadIdx.x;
        As[threadIdx.y][threadIdx.x] = (lr < m && ka < k) ? (transA ? A[arow * lda + ka] : A[ka * lda + arow]) : 0.0;
        Bs[threadIdx.y][threadIdx.x] = (tj < n && kb < k) ? (transB ? B[kb * ldb + bcol] : B[bcol * ldb + kb]) : 0.0;
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < 32; ++kk) {
            const double a = As[threadIdx.x][kk], b = Bs[threadIdx.y][kk];
            const double p = a * b;
            const double pe = fma(a, b, -p);
            dd_accumulate(p, pe, hi, lo);
        }
        __syncthreads();
    }
    if (ti < m && tj < n) {
        C1[tj * ldc + ti] = hi;
        C2[tj * ldc + ti] = lo;
    }
}

}  // namespace

extern "C" {

// dtype: 0 f32, 1 f64, 2 c32, 3 c64 (same tags as gemmul8_dtype_t); count = number of elements
int gemmul8_aux_phi_matrix(int dtype, size_t count, void *out, double phi, unsigned long long seed, void *stream) {
    if (count == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned grid = (unsigned)((count + 255) / 256);
    switch (dtype) {
        case 0: phi_matrix_kernel<float, false><<<grid, 256, 0, st>>>(count, static_cast<float *>(out), (float)phi, seed); break;
        case 1: phi_matrix_kernel<double, false><<<grid, 256, 0, st>>>(count, static_cast<double *>(out), phi, seed); break;
        case 2: phi_matrix_kernel<float, true><<<grid, 256, 0, st>>>(count, static_cast<float *>(out), (float)phi, seed); break;
        case 3: phi_matrix_kernel<double, true><<<grid, 256, 0, st>>>(count, static_cast<double *>(out), phi, seed); break;
        default: return 2;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : 3;
}

// (C1, C2)[i, j] = sum_k op(A)[rows[i], k] * op(B)[k, cols[j]]; rows / cols may be NULL (identity)
int gemmul8_aux_dd_gemm(size_t m, size_t n, size_t k, const double *A, size_t lda, int transA, const double *B, size_t ldb,
                        int transB, const int *rows, const int *cols, double *C1, double *C2, size_t ldc, void *stream) {
    if (m == 0 || n == 0) return 0;
    dim3 block(32, 32), grid((unsigned)((m + 31) / 32), (unsigned)((n + 31) / 32));
    dd_gemm_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(m, n, k, A, lda, transA, B, ldb, transB, rows, cols, C1, C2, ldc);
    return cudaGetLastError() == cudaSuccess ? 0 : 3;
}

}  // extern "C"
