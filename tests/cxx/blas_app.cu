// An application that knows nothing about this repository: plain cuBLAS calls (column-major, host
// alpha / beta, and one call in CUBLAS_POINTER_MODE_DEVICE).  tests/test_interposer.py runs it with and without
// LD_PRELOAD=libgemmul8_b200_blas.so and compares the printed samples of C.
#include <cublas_v2.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <vector>

template <typename T> static void fill(std::vector<T> &v, unsigned seed) {
    unsigned long long s = seed * 2654435761ull + 12345;
    double *p = reinterpret_cast<double *>(v.data());   // T is double or cuDoubleComplex
    for (size_t i = 0; i < v.size() * sizeof(T) / sizeof(double); ++i) {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        p[i] = ((double)(s >> 11) / 9007199254740992.0) - 0.5;
    }
}

int main() {
    cublasHandle_t h;
    cublasCreate(&h);
    cudaStream_t st;
    cudaStreamCreate(&st);
    cublasSetStream(h, st);
    {   // DGEMM 1024 x 768 x 2048, C = 1.5 A B - 0.5 C
        const int m = 1024, n = 768, k = 2048;
        std::vector<double> A((size_t)m * k), B((size_t)k * n), C((size_t)m * n);
        fill(A, 1); fill(B, 2); fill(C, 3);
        double *dA, *dB, *dC;
        cudaMalloc(&dA, A.size() * 8); cudaMalloc(&dB, B.size() * 8); cudaMalloc(&dC, C.size() * 8);
        cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice);
        cudaMemcpy(dC, C.data(), C.size() * 8, cudaMemcpyHostToDevice);
        const double alpha = 1.5, beta = -0.5;
        cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, m, n, k, &alpha, dA, m, dB, k, &beta, dC, m);
        cudaStreamSynchronize(st);
        cudaMemcpy(C.data(), dC, C.size() * 8, cudaMemcpyDeviceToHost);
        for (size_t i = 0; i < C.size(); i += 99991) printf("D %zu %.17g\n", i, C[i]);
    }
    {   // ZGEMM 512 x 384 x 1024 with A^H
        const int m = 512, n = 384, k = 1024;
        std::vector<cuDoubleComplex> A((size_t)k * m), B((size_t)k * n), C((size_t)m * n);
        fill(A, 4); fill(B, 5);
        cuDoubleComplex *dA, *dB, *dC;
        cudaMalloc(&dA, A.size() * 16); cudaMalloc(&dB, B.size() * 16); cudaMalloc(&dC, C.size() * 16);
        cudaMemcpy(dA, A.data(), A.size() * 16, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, B.data(), B.size() * 16, cudaMemcpyHostToDevice);
        const cuDoubleComplex alpha = make_cuDoubleComplex(1.0, 0.0), beta = make_cuDoubleComplex(0.0, 0.0);
        cublasZgemm(h, CUBLAS_OP_C, CUBLAS_OP_N, m, n, k, &alpha, dA, k, dB, k, &beta, dC, m);
        cudaStreamSynchronize(st);
        cudaMemcpy(C.data(), dC, C.size() * 16, cudaMemcpyDeviceToHost);
        for (size_t i = 0; i < C.size(); i += 39989) printf("Z %zu %.17g %.17g\n", i, C[i].x, C[i].y);
    }
    {   // DGEMM with alpha / beta in DEVICE memory (CUBLAS_POINTER_MODE_DEVICE), C = -2 A^T B + 0.25 C
        const int m = 768, n = 1024, k = 2048;
        std::vector<double> A((size_t)k * m), B((size_t)k * n), C((size_t)m * n);
        fill(A, 6); fill(B, 7); fill(C, 8);
        double *dA, *dB, *dC, *dS;
        cudaMalloc(&dA, A.size() * 8); cudaMalloc(&dB, B.size() * 8); cudaMalloc(&dC, C.size() * 8); cudaMalloc(&dS, 16);
        cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice);
        cudaMemcpy(dC, C.data(), C.size() * 8, cudaMemcpyHostToDevice);
        const double scal[2] = {-2.0, 0.25};
        cudaMemcpy(dS, scal, 16, cudaMemcpyHostToDevice);
        cublasSetPointerMode(h, CUBLAS_POINTER_MODE_DEVICE);
        cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, m, n, k, dS, dA, k, dB, k, dS + 1, dC, m);
        cublasSetPointerMode(h, CUBLAS_POINTER_MODE_HOST);
        cudaStreamSynchronize(st);
        cudaMemcpy(C.data(), dC, C.size() * 8, cudaMemcpyDeviceToHost);
        for (size_t i = 0; i < C.size(); i += 99991) printf("P %zu %.17g\n", i, C[i]);
    }
    {   // a tiny DGEMM that must stay with cuBLAS (below GEMMUL8_MIN_MNK)
        const int m = 8, n = 8, k = 8;
        std::vector<double> A(64, 1.0), B(64, 2.0), C(64, 0.0);
        double *dA, *dB, *dC;
        cudaMalloc(&dA, 512); cudaMalloc(&dB, 512); cudaMalloc(&dC, 512);
        cudaMemcpy(dA, A.data(), 512, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, B.data(), 512, cudaMemcpyHostToDevice);
        const double one = 1.0, zero = 0.0;
        cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, m, n, k, &one, dA, m, dB, k, &zero, dC, m);
        cudaStreamSynchronize(st);
        cudaMemcpy(C.data(), dC, 512, cudaMemcpyDeviceToHost);
        printf("T 0 %.17g\n", C[0]);
    }
    printf("cuda status %d\n", (int)cudaGetLastError());
    return 0;
}
