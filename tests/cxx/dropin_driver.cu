// A caller written against the reference's public API only (gemmul8.hpp): what a user of
// ptrkgtsch/mixed-GEMMul8 has in their code base (cf. GEMMul8/testing/test_double.cu:120-170).
// Built against include/gemmul8.hpp and linked with libgemmul8_b200.so by tests/test_cxx_dropin.py.
//   dropin_driver worksize          -> prints workSize(1024,1024,1024,14) (no GPU needed)
//   dropin_driver run               -> DGEMM / SGEMM / mixed emulation vs cuBLAS on the GPU
#include "gemmul8.hpp"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

template <typename T> static void fill(std::vector<T> &v, unsigned seed) {
    unsigned long long s = seed * 2654435761ull + 12345;
    for (auto &x : v) {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        x = (T)(((double)(s >> 11) / 9007199254740992.0) - 0.5);
    }
}

template <typename TA, typename TB, typename TC>
static double run_case(cublasHandle_t handle, size_t m, size_t n, size_t k, unsigned N, bool fast) {
    std::vector<TA> hA(m * k);
    std::vector<TB> hB(k * n);
    fill(hA, 1); fill(hB, 2);
    TA *dA; TB *dB; TC *dC; double *dA64, *dB64, *dC64; void *work;
    cudaMalloc(&dA, sizeof(TA) * m * k); cudaMalloc(&dB, sizeof(TB) * k * n); cudaMalloc(&dC, sizeof(TC) * m * n);
    cudaMalloc(&dA64, 8 * m * k); cudaMalloc(&dB64, 8 * k * n); cudaMalloc(&dC64, 8 * m * n);
    cudaMalloc(&work, gemmul8::workSize(m, n, k, N));
    cudaMemcpy(dA, hA.data(), sizeof(TA) * m * k, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), sizeof(TB) * k * n, cudaMemcpyHostToDevice);
    std::vector<double> a64(hA.begin(), hA.end()), b64(hB.begin(), hB.end());
    cudaMemcpy(dA64, a64.data(), 8 * m * k, cudaMemcpyHostToDevice);
    cudaMemcpy(dB64, b64.data(), 8 * k * n, cudaMemcpyHostToDevice);
    const TC alpha = 1, beta = 0;
    const double one = 1.0, zero = 0.0;
    std::vector<double> times =
        gemmul8::gemm<TA, TB, TC>(handle, GPUBLAS_OP_N, GPUBLAS_OP_N, m, n, k, &alpha, dA, m, dB, k, &beta, dC, m, N, fast, work);
    cublasDgemm(handle, CUBLAS_OP_N, CUBLAS_OP_N, (int)m, (int)n, (int)k, &one, dA64, (int)m, dB64, (int)k, &zero, dC64, (int)m);
    std::vector<TC> c(m * n);
    std::vector<double> c64(m * n);
    cudaMemcpy(c.data(), dC, sizeof(TC) * m * n, cudaMemcpyDeviceToHost);
    cudaMemcpy(c64.data(), dC64, 8 * m * n, cudaMemcpyDeviceToHost);
    double err = 0, scale = 0;
    for (size_t i = 0; i < m * n; ++i) { err = fmax(err, fabs((double)c[i] - c64[i])); scale = fmax(scale, fabs(c64[i])); }
    if (times.size() != 4 || !(times[0] > 0.0) || !(times[1] > 0.0)) return 1e300;   // 4 phase times in ns
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dA64); cudaFree(dB64); cudaFree(dC64); cudaFree(work);
    return err / scale;
}

// cuDoubleComplex through the three computeType_t values, against cublasZgemm
static double run_complex(cublasHandle_t handle, gemmul8::computeType_t ct) {
    const size_t m = 200, n = 160, k = 256;
    const unsigned N = 14;
    std::vector<cuDoubleComplex> hA(m * k), hB(k * n), c(m * n), cz(m * n);
    std::vector<double> ra(2 * m * k), rb(2 * k * n);
    fill(ra, 11); fill(rb, 12);
    for (size_t i = 0; i < m * k; ++i) hA[i] = make_cuDoubleComplex(ra[2 * i], ra[2 * i + 1]);
    for (size_t i = 0; i < k * n; ++i) hB[i] = make_cuDoubleComplex(rb[2 * i], rb[2 * i + 1]);
    cuDoubleComplex *dA, *dB, *dC, *dZ;
    void *work;
    cudaMalloc(&dA, 16 * m * k); cudaMalloc(&dB, 16 * k * n); cudaMalloc(&dC, 16 * m * n); cudaMalloc(&dZ, 16 * m * n);
    cudaMalloc(&work, gemmul8::workSize(m, n, k, N, ct));
    cudaMemcpy(dA, hA.data(), 16 * m * k, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), 16 * k * n, cudaMemcpyHostToDevice);
    const cuDoubleComplex one = make_cuDoubleComplex(1, 0), zero = make_cuDoubleComplex(0, 0);
    gemmul8::gemm<gpuDoubleComplex>(handle, GPUBLAS_OP_N, GPUBLAS_OP_N, m, n, k, &one, dA, m, dB, k, &zero, dC, m, N, true, work, ct);
    cublasZgemm(handle, CUBLAS_OP_N, CUBLAS_OP_N, (int)m, (int)n, (int)k, &one, dA, (int)m, dB, (int)k, &zero, dZ, (int)m);
    cudaMemcpy(c.data(), dC, 16 * m * n, cudaMemcpyDeviceToHost);
    cudaMemcpy(cz.data(), dZ, 16 * m * n, cudaMemcpyDeviceToHost);
    double err = 0, scale = 0;
    for (size_t i = 0; i < m * n; ++i) {
        err = fmax(err, fmax(fabs(c[i].x - cz[i].x), fabs(c[i].y - cz[i].y)));
        scale = fmax(scale, fmax(fabs(cz[i].x), fabs(cz[i].y)));
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dZ); cudaFree(work);
    return err / scale;
}

int main(int argc, char **argv) {
    if (argc > 1 && !strcmp(argv[1], "worksize")) {
        printf("%zu %zu\n", gemmul8::workSize(1024, 1024, 1024, 14),
               gemmul8::workSize(512, 256, 128, 9, gemmul8::COMPLEX_KARATSUBA_MULT));
        return 0;
    }
    cublasHandle_t handle;
    if (cublasCreate(&handle) != CUBLAS_STATUS_SUCCESS) { printf("no cublas\n"); return 2; }
    cudaStream_t st;
    cudaStreamCreate(&st);
    cublasSetStream(handle, st);   // the shim runs on the handle's stream
    const double e_d  = run_case<double, double, double>(handle, 300, 200, 400, 14, true);
    const double e_da = run_case<double, double, double>(handle, 300, 200, 400, 15, false);
    const double e_s  = run_case<float, float, float>(handle, 257, 129, 300, 7, true);
    const double e_m  = run_case<double, float, double>(handle, 128, 128, 256, 12, true);
    const double e_zb = run_complex(handle, gemmul8::COMPLEX_BIG_MATRIX_ENCODE);
    const double e_zc = run_complex(handle, gemmul8::COMPLEX_CLASSIC_MULT);
    const double e_zk = run_complex(handle, gemmul8::COMPLEX_KARATSUBA_MULT);
    printf("relerr dgemm-fast %.3e dgemm-accurate %.3e sgemm %.3e mixed %.3e zgemm big %.3e classic %.3e karatsuba %.3e\n", e_d, e_da, e_s,
           e_m, e_zb, e_zc, e_zk);
    // unsupported computeType for real types: message on stderr, zero timers (GEMMul8/src/gemmul8.cu:174-177)
    double *dummy; void *work;
    cudaMalloc(&dummy, 8 * 64); cudaMalloc(&work, gemmul8::workSize(8, 8, 8, 4));
    const double one = 1, zero = 0;
    std::vector<double> t = gemmul8::gemm<double>(handle, GPUBLAS_OP_N, GPUBLAS_OP_N, 8, 8, 8, &one, dummy, 8, dummy, 8, &zero, dummy, 8, 4,
                                                 true, work, gemmul8::COMPLEX_CLASSIC_MULT);
    const bool zeros = t.size() == 4 && t[0] == 0 && t[1] == 0 && t[2] == 0 && t[3] == 0;
    const bool ok = e_d < 1e-12 && e_da < 1e-12 && e_s < 1e-5 && e_m < 1e-6 && e_zb < 1e-12 && e_zc < 1e-12 && e_zk < 1e-12 && zeros;
    printf(ok ? "DROPIN OK\n" : "DROPIN FAILED\n");
    return ok ? 0 : 1;
}
