// Complex operands: residue encoding into the reference's two complex workspace layouts, the
// Karatsuba slice sum, the accurate-mode bound slices and the complex CRT.
//
// Layouts (both exactly as the reference writes them, so parity tests compare them in place):
//   PLANES  (COMPLEX_CLASSIC_MULT / COMPLEX_KARATSUBA_MULT): separate real and imaginary slice
//           stacks A8i_real / A8i_imag [N][m_pad][lda8i], lda8i = ceil16(k)
//           reference: scalingA_kara / scalingB_kara (+ _conj), GEMMul8/src/scaling.hpp:840-923,
//           :1006-1089, :1232-1319, :1411-1498
//   BIG     (COMPLEX_BIG_MATRIX_ENCODE): A^ = [Pr -Pi; Pi Pr] (2m x 2k, K-major rows, P = op(A)),
//           B^ = [Qr; Qi] (2k x n, Q = op(B)), lda8i = ceil16(2k); the imaginary half starts at
//           byte k of a row (NOT at a padded offset)
//           reference: scalingA_bigmatrix / scalingB_bigmatrix (+ _minusTR / _minusBL),
//           GEMMul8/src/scaling.hpp:753-838, :925-1004, :1150-1230, :1321-1409
// A conjugate-transposed operand negates the residues of its imaginary part (the symmetric
// residue is odd, and for modulus 256 both signs of 128 wrap to -128, so "negate the residue" and
// "residue of the negated value" are the same byte).
//
// The shifts come from the kernels in oz_scale.cu (complex element types are handled there: amax
// over |re|, |im|, sum of squares over both, GEMMul8/src/scaling.hpp:155-213).
#include "oz_common.cuh"
#include "oz_crt.cuh"
#include "oz_residue.cuh"

namespace oz {
namespace {

template <typename T> struct CReal;
template <> struct CReal<float2> { using type = float; };
template <> struct CReal<double2> { using type = double; };

enum : int { LAYOUT_PLANES = 0, LAYOUT_BIG_A = 1, LAYOUT_BIG_B = 2 };

struct CplxSink {
    int layout;
    int8_t *out_re;    // PLANES: real stack;  BIG: the big matrix stack
    int8_t *out_im;    // PLANES: imaginary stack
    size_t ld8i, inc;  // row stride, slice stride
    size_t k;          // logical inner length: offset of the second half of a BIG row
    size_t nvec;       // BIG_A: row offset of the second block row
};

template <int G> __device__ __forceinline__ void store_group(int8_t *dst, const uint32_t *w, int count) {
    if (count <= 0) return;
    if (count == G && (reinterpret_cast<uintptr_t>(dst) & (G - 1)) == 0) {
        if constexpr (G == 4) *reinterpret_cast<uint32_t *>(dst) = w[0];
        else if constexpr (G == 8) *reinterpret_cast<uint2 *>(dst) = make_uint2(w[0], w[1]);
        else *reinterpret_cast<uint4 *>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
        for (int e = 0; e < count; ++e) dst[e] = (int8_t)((w[e >> 2] >> (8 * (e & 3))) & 0xffu);
    }
}

// G consecutive positions i0 .. i0+G-1 of vector `vec`, modulus j.  re / im hold G scaled, truncated
// values each (zeros beyond k); `sgn` = -1 for a conjugated operand.
template <typename R, int G, bool SPLIT>
__device__ __forceinline__ void encode_group(const CplxSink &s, size_t vec, size_t i0, const R (&re)[G], const R (&im)[G],
                                             int sgn, unsigned num_moduli, bool ref_chain) {
    const int first  = (int)min((size_t)G, s.k > i0 ? s.k - i0 : (size_t)0);                          // positions < k
    const int second = (int)min((size_t)G, s.ld8i - s.k > i0 ? s.ld8i - s.k - i0 : (size_t)0);        // positions < ld8i - k
    R both[2 * G];
#pragma unroll
    for (int e = 0; e < G; ++e) { both[e] = re[e]; both[G + e] = im[e]; }
    residues_of<2 * G, SPLIT>(both, num_moduli, ref_chain, [&](unsigned j, const int (&q)[2 * G]) {
        int rr[G], ri[G];
#pragma unroll
        for (int e = 0; e < G; ++e) { rr[e] = q[e]; ri[e] = sgn * q[G + e]; }
        uint32_t pr[G / 4], pi[G / 4];
#pragma unroll
        for (int q = 0; q < G / 4; ++q) {
            pr[q] = pack4(rr[4 * q], rr[4 * q + 1], rr[4 * q + 2], rr[4 * q + 3]);
            pi[q] = pack4(ri[4 * q], ri[4 * q + 1], ri[4 * q + 2], ri[4 * q + 3]);
        }
        if (s.layout == LAYOUT_PLANES) {
            const size_t off = (size_t)j * s.inc + vec * s.ld8i + i0;
            store_group<G>(s.out_re + off, pr, G);
            store_group<G>(s.out_im + off, pi, G);
        } else {
            int8_t *row = s.out_re + (size_t)j * s.inc + vec * s.ld8i;
            if (s.layout == LAYOUT_BIG_B) {            // [ Qr | Qi ]
                store_group<G>(row + i0, pr, first);
                store_group<G>(row + s.k + i0, pi, second);
            } else {                                   // [ Pr | -Pi ] and, nvec rows below, [ Pi | Pr ]
                uint32_t pn[G / 4];
#pragma unroll
                for (int q = 0; q < G / 4; ++q)
                    pn[q] = pack4(-ri[4 * q], -ri[4 * q + 1], -ri[4 * q + 2], -ri[4 * q + 3]);
                int8_t *row2 = row + s.nvec * s.ld8i;
                store_group<G>(row + i0, pr, first);
                store_group<G>(row + s.k + i0, pn, second);
                store_group<G>(row2 + i0, pi, first);
                store_group<G>(row2 + s.k + i0, pr, second);
            }
        }
    });
}

// contiguous vectors: one thread = 4 consecutive elements of one vector; grid = (ceil(ld8i/4/256), nvec)
template <typename T, bool SPLIT>
__global__ void __launch_bounds__(256) encode_cplx_contig_kernel(const T *__restrict__ X, size_t ld, size_t len,
                                                                 const int16_t *__restrict__ sft_neg, unsigned num_moduli,
                                                                 CplxSink sink, int sgn, bool ref_chain) {
    using R = typename CReal<T>::type;
    const size_t vec = blockIdx.y;
    const size_t i0  = ((size_t)blockIdx.x * 256 + threadIdx.x) * 4;
    const size_t lim = sink.layout == LAYOUT_PLANES ? sink.ld8i : sink.ld8i - sink.k;
    if (i0 >= lim) return;
    const Pow2<R> scale(-(int)sft_neg[vec]);
    const T *__restrict__ p = X + vec * ld;
    R re[4], im[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        if (i0 + e < len) { const T x = p[i0 + e]; re[e] = scale(x.x); im[e] = scale(x.y); }
        else { re[e] = R(0); im[e] = R(0); }
    }
    encode_group<R, 4, SPLIT>(sink, vec, i0, re, im, sgn, num_moduli, ref_chain);
}

// strided vectors: tile of 32 vectors x 64 k through shared memory (lane == vector on the way in,
// 8 consecutive k per thread on the way out); grid = (ceil(ld8i/64), ceil(nvec/32)), 256 threads
template <typename T, bool SPLIT>
__global__ void __launch_bounds__(256, 2) encode_cplx_strided_kernel(const T *__restrict__ X, size_t ld, size_t nvec, size_t len,
                                                                  const int16_t *__restrict__ sft_neg, unsigned num_moduli,
                                                                  CplxSink sink, int sgn, bool ref_chain) {
    using R = typename CReal<T>::type;
    __shared__ R tile_re[64 * 32];
    __shared__ R tile_im[64 * 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t v0 = (size_t)blockIdx.y * 32, k0 = (size_t)blockIdx.x * 64;
    const size_t lim = sink.layout == LAYOUT_PLANES ? sink.ld8i : sink.ld8i - sink.k;
    if (k0 >= lim) return;
    {
        const size_t vec = v0 + lane;
        const bool active = vec < nvec;
        const Pow2<R> scale(active ? -(int)sft_neg[vec] : 0);
        const T *__restrict__ p = X + (active ? vec : 0);
#pragma unroll 8
        for (int kk = warp; kk < 64; kk += 8) {
            const size_t k = k0 + kk;
            R xr = 0, xi = 0;
            if (active && k < len) { const T x = p[k * ld]; xr = scale(x.x); xi = scale(x.y); }
            const int slot = kk * 32 + ((lane + 4 * (kk >> 3)) & 31);
            tile_re[slot] = xr;
            tile_im[slot] = xi;
        }
    }
    __syncthreads();
    const int r = threadIdx.x >> 3, g = threadIdx.x & 7;
    const size_t vec = v0 + r, i0 = k0 + 8 * g;
    if (vec >= nvec || i0 >= lim) return;
    R re[8], im[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int slot = (8 * g + e) * 32 + ((r + 4 * g) & 31);
        re[e] = tile_re[slot];
        im[e] = tile_im[slot];
    }
    encode_group<R, 8, SPLIT>(sink, vec, i0, re, im, sgn, num_moduli, ref_chain);
}

template <typename T>
cudaError_t run_encode_complex(bool strided, const void *X, size_t ld, size_t nvec, size_t len, const int16_t *sft_neg,
                               unsigned N, const CplxSink &sink, int sgn, cudaStream_t st) {
    if (nvec == 0) return cudaSuccess;
    const T *x = static_cast<const T *>(X);
    if (strided) {
        dim3 grid((unsigned)((sink.ld8i + 63) / 64), (unsigned)((nvec + 31) / 32));
        (N >= 16 ? encode_cplx_strided_kernel<T, true> : encode_cplx_strided_kernel<T, false>)<<<grid, 256, 0, st>>>(x, ld, nvec, len, sft_neg, N, sink, sgn, encode_reference_chain());
        count_launch();
    } else {
        for (size_t v0 = 0; v0 < nvec; v0 += 65535) {
            const size_t nv = nvec - v0 < 65535 ? nvec - v0 : 65535;
            CplxSink s = sink;
            s.out_re += v0 * sink.ld8i;
            if (s.out_im) s.out_im += v0 * sink.ld8i;
            dim3 grid((unsigned)((sink.ld8i / 4 + 255) / 256), (unsigned)nv);
            (N >= 16 ? encode_cplx_contig_kernel<T, true> : encode_cplx_contig_kernel<T, false>)<<<grid, 256, 0, st>>>(x + v0 * ld, ld, len, sft_neg + v0, N, s, sgn, encode_reference_chain());
            count_launch();
        }
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Karatsuba: G = (X_real + X_imag) mod m_j back to an int8 representative, in place in X_real.
// reference: add_int8_mat_256_kernel / add_int8_mat_not256_kernel, GEMMul8/src/mat_utils.hpp:6-67
// (same integer steps, so the representative -- in [-(m+1)/2, (m-3)/2] for odd m -- is identical).
// One launch for all moduli: grid.y = modulus.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int add_mod_sym(int a, int b, int m, int invm) {
    int t = a + b;
    t -= __mulhi(t, invm) * m;
    t -= (t >= m / 2) ? m : 0;
    t += (t < -(m / 2)) ? m : 0;
    return t;
}
__global__ void __launch_bounds__(256) add_int8_slices_kernel(size_t words, size_t slice_stride, int8_t *X_real,
                                                              const int8_t *__restrict__ X_imag) {
    const unsigned j = blockIdx.y;
    const int m      = dev_tab::OZ_MOD[j];
    const int invm   = (int)(4294967296ull / (unsigned)m);
    uint32_t *xr       = reinterpret_cast<uint32_t *>(X_real + (size_t)j * slice_stride);
    const uint32_t *xi = reinterpret_cast<const uint32_t *>(X_imag + (size_t)j * slice_stride);
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < words; i += (size_t)gridDim.x * 256) {
        const uint32_t a = xr[i], b = xi[i];
        int r[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int ae = (int)(int8_t)(a >> (8 * e)), be = (int)(int8_t)(b >> (8 * e));
            r[e] = (j == 0) ? ae + be : add_mod_sym(ae, be, m, invm);   // modulus 256: the int8 cast wraps
        }
        xr[i] = pack4(r[0], r[1], r[2], r[3]);
    }
}

// ---------------------------------------------------------------------------------------------
// accurate mode, complex: amax over |re|, |im| -> sft0 = 5 - ilogb(amax); bound slices
// ceil(|re| 2^sft0), ceil(|im| 2^sft0) in the PLANES or BIG layout.
// reference: extract_A8i_kernel_{bigmatrix,kara}* / extract_B8i_kernel_*, scaling.hpp:1943-2532.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) amax_cplx_contig_kernel(const T *__restrict__ X, size_t ld, size_t len,
                                                               int16_t *__restrict__ sft0) {
    using R = typename CReal<T>::type;
    __shared__ R s_max[8];
    const T *__restrict__ p = X + (size_t)blockIdx.x * ld;
    R amax = 0;
    for (size_t i = threadIdx.x; i < len; i += 256) { const T x = p[i]; amax = fmax(amax, fmax(fabs(x.x), fabs(x.y))); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = amax;
    __syncthreads();
    if (threadIdx.x == 0) {
        R a = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) a = fmax(a, s_max[w]);
        int e;
        if constexpr (sizeof(R) == 8) e = ilogb(a); else e = ilogbf(a);
        sft0[blockIdx.x] = (a == R(0)) ? (int16_t)0 : (int16_t)(5 - e);
    }
}
template <typename T>
__global__ void __launch_bounds__(512) amax_cplx_strided_kernel(const T *__restrict__ X, size_t ld, size_t nvec, size_t len,
                                                                int16_t *__restrict__ sft0) {
    using R = typename CReal<T>::type;
    __shared__ R s_max[16 * 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t vec = (size_t)blockIdx.x * 32 + lane;
    R amax = 0;
    if (vec < nvec) {
        const T *__restrict__ p = X + vec;
        for (size_t i = warp; i < len; i += 16) { const T x = p[i * ld]; amax = fmax(amax, fmax(fabs(x.x), fabs(x.y))); }
    }
    s_max[warp * 32 + lane] = amax;
    __syncthreads();
    if (warp == 0 && vec < nvec) {
        R a = 0;
#pragma unroll
        for (int w = 0; w < 16; ++w) a = fmax(a, s_max[w * 32 + lane]);
        int e;
        if constexpr (sizeof(R) == 8) e = ilogb(a); else e = ilogbf(a);
        sft0[vec] = (a == R(0)) ? (int16_t)0 : (int16_t)(5 - e);
    }
}
__device__ __forceinline__ int bound_int_c(double x, int sft) { return __double2int_ru(scalbn(fabs(x), sft)); }
__device__ __forceinline__ int bound_int_c(float x, int sft) { return __float2int_ru(scalbnf(fabsf(x), sft)); }

// one thread = one element (both parts) of one vector; simple on purpose: the bound pass is a small
// fraction of the accurate mode (one extra int8 GEMM follows it).  Always the BIG layout: the bound
// product needs Re = |Pr||Qr| - |Pi||Qi| and Im = |Pi||Qr| + |Pr||Qi| element by element (that is what
// the reference takes row / column maxima of, in all three compute types: scaling.hpp:3222, :3337-3344),
// and the sign pattern of the big matrix turns both into ONE product.  `sgn` = -1 for a conjugated
// operand: the reference keeps the residue layout's signs on the absolute values (scaling.hpp:2151-2213).
template <typename T>
__global__ void __launch_bounds__(256) bound_cplx_kernel(const T *__restrict__ X, size_t ld, size_t nvec, size_t len, bool strided,
                                                         const int16_t *__restrict__ sft0, CplxSink sink, int sgn) {
    const size_t lim = sink.ld8i - sink.k;
    size_t vec, i;
    if (strided) { vec = (size_t)blockIdx.x * 32 + (threadIdx.x & 31); i = (size_t)blockIdx.y * 8 + (threadIdx.x >> 5); }
    else         { vec = blockIdx.y; i = (size_t)blockIdx.x * 256 + threadIdx.x; }
    if (vec >= nvec || i >= lim) return;
    int br = 0, bi = 0;
    if (i < len) {
        const T x   = strided ? X[vec + i * ld] : X[vec * ld + i];
        const int s = sft0[vec];
        br = bound_int_c(x.x, s);
        bi = sgn * bound_int_c(x.y, s);
    }
    int8_t *row = sink.out_re + vec * sink.ld8i;
    if (sink.layout == LAYOUT_BIG_B) {            // [ |Qr| | +-|Qi| ]
        if (i < sink.k) row[i] = (int8_t)br;
        row[sink.k + i] = (int8_t)bi;
    } else {                                      // [ |Pr| | -+|Pi| ] and, nvec rows below, [ +-|Pi| | |Pr| ]
        int8_t *row2 = row + sink.nvec * sink.ld8i;
        if (i < sink.k) { row[i] = (int8_t)br; row2[i] = (int8_t)bi; }
        row[sink.k + i]  = (int8_t)(-bi);
        row2[sink.k + i] = (int8_t)br;
    }
}

// ---------------------------------------------------------------------------------------------
// complex CRT: residues of Re and Im (two stacks with a common leading dimension) -> C
// reference: inverse_scaling_{1,2}_base_{bigmatrix,kara}, GEMMul8/src/inverse_scaling.hpp:64-136,
// :174-262; alpha/beta forms :294-820 (Tfma / CMul / CAdd on cuComplex)
// ---------------------------------------------------------------------------------------------
template <typename T2> struct CplxOps;
template <> struct CplxOps<double2> {
    using R = double;
    static __device__ __forceinline__ double2 make(double re, double im) { return make_double2(re, im); }
    static __device__ __forceinline__ double2 mul(double2 a, double2 b) { return cuCmul(a, b); }
    static __device__ __forceinline__ double2 add(double2 a, double2 b) { return cuCadd(a, b); }
    static __device__ __forceinline__ double2 fma(double2 a, double2 b, double2 c) { return cuCfma(a, b, c); }
};
template <> struct CplxOps<float2> {
    using R = float;
    static __device__ __forceinline__ float2 make(double re, double im) { return make_float2(__double2float_rn(re), __double2float_rn(im)); }
    static __device__ __forceinline__ float2 mul(float2 a, float2 b) { return cuCmulf(a, b); }
    static __device__ __forceinline__ float2 add(float2 a, float2 b) { return cuCaddf(a, b); }
    static __device__ __forceinline__ float2 fma(float2 a, float2 b, float2 c) { return cuCfmaf(a, b, c); }
};

template <typename T2>
__device__ __forceinline__ T2 combine_c(int mode, T2 alpha, T2 beta, T2 c, const T2 *cptr) {
    using O = CplxOps<T2>;
    switch (mode) {
        case AB_10: return c;
        case AB_11: return O::add(*cptr, c);
        case AB_1B: return O::fma(beta, *cptr, c);
        case AB_A0: return O::mul(alpha, c);
        case AB_A1: return O::fma(alpha, c, *cptr);
        default:    return O::fma(beta, *cptr, O::mul(alpha, c));
    }
}

template <typename T2> __host__ __device__ inline int alpha_beta_mode_c(T2 alpha, T2 beta) {
    const bool a1 = alpha.x == 1 && alpha.y == 0, b0 = beta.x == 0 && beta.y == 0, b1 = beta.x == 1 && beta.y == 0;
    if (a1) return b0 ? AB_10 : b1 ? AB_11 : AB_1B;
    return b0 ? AB_A0 : b1 ? AB_A1 : AB_AB;
}

__device__ __forceinline__ uint32_t load4(const uint8_t *p) {
    if ((reinterpret_cast<uintptr_t>(p) & 3) == 0) return *reinterpret_cast<const uint32_t *>(p);
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

template <typename T2, bool SPLIT, int N>
__global__ void __launch_bounds__(256) crt_cplx_kernel(size_t m, size_t n, const uint8_t *__restrict__ Cre,
                                                       const uint8_t *__restrict__ Cim, size_t ldc8u, size_t sizeC,
                                                       T2 *__restrict__ C, size_t ldc, const int16_t *__restrict__ sftA,
                                                       const int16_t *__restrict__ sftB, int mode, T2 alpha, T2 beta,
                                                       const T2 *__restrict__ alpha_dev, const T2 *__restrict__ beta_dev) {
    if (alpha_dev != nullptr) {   // device-resident scalars
        alpha = *alpha_dev; beta = *beta_dev;
        mode  = alpha_beta_mode_c(alpha, beta);
    }
    const size_t row0 = ((size_t)blockIdx.x * 64 + threadIdx.x) * 4;
    const size_t col  = (size_t)blockIdx.y * 4 + threadIdx.y;
    if (row0 >= m || col >= n) return;
    // rows row0 .. row0+3 may run past m (then into the next block row / padding of the stack: still
    // inside the buffer because every stack is followed by more workspace); such lanes are not stored
    uint32_t rr[N], ri[N];
    const uint8_t *sr = Cre + col * ldc8u + row0, *si = Cim + col * ldc8u + row0;
#pragma unroll
    for (int j = 0; j < N; ++j) { rr[j] = load4(sr + (size_t)j * sizeC); ri[j] = load4(si + (size_t)j * sizeC); }
    double a1[4] = {0, 0, 0, 0}, a2[4] = {0, 0, 0, 0}, b1[4] = {0, 0, 0, 0}, b2[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < N; ++j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            crt_step<SPLIT>(N, j, __byte_perm(rr[j], 0, 0x4440 + e), a1[e], a2[e]);
            crt_step<SPLIT>(N, j, __byte_perm(ri[j], 0, 0x4440 + e), b1[e], b2[e]);
        }
    }
    const int sb = sftB[col];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const size_t row = row0 + e;
        if (row < m) {
            const int s     = (int)sftA[row] + sb;
            const double re = scale_pow2(crt_finish<SPLIT>(N, a1[e], a2[e]), s);
            const double im = scale_pow2(crt_finish<SPLIT>(N, b1[e], b2[e]), s);
            T2 *cptr = C + col * ldc + row;
            *cptr    = combine_c<T2>(mode, alpha, beta, CplxOps<T2>::make(re, im), cptr);
        }
    }
}

template <typename T2, bool SPLIT, int N>
void launch_crt_cplx_one(dim3 grid, dim3 block, cudaStream_t st, size_t m, size_t n, const uint8_t *Cre, const uint8_t *Cim,
                         size_t ldc8u, size_t sizeC, void *C, size_t ldc, const int16_t *sftA, const int16_t *sftB, int mode,
                         T2 alpha, T2 beta, const T2 *alpha_dev, const T2 *beta_dev) {
    crt_cplx_kernel<T2, SPLIT, N><<<grid, block, 0, st>>>(m, n, Cre, Cim, ldc8u, sizeC, static_cast<T2 *>(C), ldc, sftA, sftB,
                                                          mode, alpha, beta, alpha_dev, beta_dev);
}

template <typename T2>
cudaError_t run_crt_cplx(bool split, unsigned N, size_t m, size_t n, const uint8_t *Cre, const uint8_t *Cim, size_t ldc8u,
                         size_t sizeC, void *C, size_t ldc, const int16_t *sftA, const int16_t *sftB, const void *alpha_p,
                         const void *beta_p, bool device_scalars, cudaStream_t st) {
    const T2 *alpha_dev = device_scalars ? static_cast<const T2 *>(alpha_p) : nullptr;
    const T2 *beta_dev  = device_scalars ? static_cast<const T2 *>(beta_p) : nullptr;
    T2 alpha{}, beta{};
    alpha.x = 1;
    if (!device_scalars) { alpha = *static_cast<const T2 *>(alpha_p); beta = *static_cast<const T2 *>(beta_p); }
    const int mode = alpha_beta_mode_c(alpha, beta);
    dim3 block(64, 4), grid((unsigned)(((m + 3) / 4 + 63) / 64), (unsigned)((n + 3) / 4));
#define OZ_CRT_CASE(NN)                                                                                                           \
    case NN:                                                                                                                      \
        if constexpr (NN >= 8 && sizeof(T2) == 16) {                                                                              \
            if (split) { launch_crt_cplx_one<T2, true, NN>(grid, block, st, m, n, Cre, Cim, ldc8u, sizeC, C, ldc, sftA, sftB, mode, alpha, beta, alpha_dev, beta_dev); break; } \
        }                                                                                                                         \
        launch_crt_cplx_one<T2, false, NN>(grid, block, st, m, n, Cre, Cim, ldc8u, sizeC, C, ldc, sftA, sftB, mode, alpha, beta, alpha_dev, beta_dev);  \
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

CplxSink make_sink(const ComplexTarget &t) {
    CplxSink s{};
    s.layout = t.layout; s.out_re = t.out_re; s.out_im = t.out_im; s.ld8i = t.ld8i; s.inc = t.inc; s.k = t.k; s.nvec = t.nvec;
    return s;
}

}  // namespace

cudaError_t launch_encode_complex(int dtype, bool strided, const void *X, size_t ld, size_t nvec, size_t len,
                                  const int16_t *sft_neg, unsigned num_moduli, const ComplexTarget &target, bool conj,
                                  cudaStream_t st) {
    if (strided && (nvec + 31) / 32 > 65535) return cudaErrorInvalidValue;
    const CplxSink s = make_sink(target);
    const int sgn    = conj ? -1 : 1;
    switch (dtype) {
        case DT_C64: return run_encode_complex<double2>(strided, X, ld, nvec, len, sft_neg, num_moduli, s, sgn, st);
        case DT_C32: return run_encode_complex<float2>(strided, X, ld, nvec, len, sft_neg, num_moduli, s, sgn, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_add_int8_slices(unsigned num_moduli, size_t slice_bytes, int8_t *X_real, const int8_t *X_imag, cudaStream_t st) {
    if (slice_bytes == 0 || num_moduli == 0) return cudaSuccess;
    const size_t words = slice_bytes / 4;   // slices are multiples of 16 bytes
    dim3 grid((unsigned)((words + 255) / 256 < 4096 ? (words + 255) / 256 : 4096), num_moduli);
    add_int8_slices_kernel<<<grid, 256, 0, st>>>(words, slice_bytes, X_real, X_imag);
    count_launch();
    return cudaGetLastError();
}

template <typename T>
static cudaError_t run_bound_complex(bool strided, const void *X, size_t ld, size_t nvec, size_t len, int16_t *sft0,
                                     const ComplexTarget &target, bool conj, cudaStream_t st) {
    if (nvec == 0) return cudaSuccess;
    if (target.layout == LAYOUT_PLANES) return cudaErrorInvalidValue;
    const T *x = static_cast<const T *>(X);
    const CplxSink s = make_sink(target);
    if (strided) amax_cplx_strided_kernel<T><<<(unsigned)((nvec + 31) / 32), 512, 0, st>>>(x, ld, nvec, len, sft0);
    else         amax_cplx_contig_kernel<T><<<(unsigned)nvec, 256, 0, st>>>(x, ld, len, sft0);
    count_launch();
    if (strided) {
        if ((s.ld8i + 7) / 8 > 65535) return cudaErrorInvalidValue;
        dim3 grid((unsigned)((nvec + 31) / 32), (unsigned)((s.ld8i + 7) / 8));
        bound_cplx_kernel<T><<<grid, 256, 0, st>>>(x, ld, nvec, len, true, sft0, s, conj ? -1 : 1);
        count_launch();
    } else {
        for (size_t v0 = 0; v0 < nvec; v0 += 65535) {
            const size_t nv = nvec - v0 < 65535 ? nvec - v0 : 65535;
            CplxSink sv = s;
            sv.out_re += v0 * s.ld8i;
            if (sv.out_im) sv.out_im += v0 * s.ld8i;
            dim3 grid((unsigned)((s.ld8i + 255) / 256), (unsigned)nv);
            bound_cplx_kernel<T><<<grid, 256, 0, st>>>(x + v0 * ld, ld, nv, len, false, sft0 + v0, sv, conj ? -1 : 1);
            count_launch();
        }
    }
    return cudaGetLastError();
}

cudaError_t launch_bound_extract_complex(int dtype, bool strided, const void *X, size_t ld, size_t nvec, size_t len,
                                         int16_t *sft_out, const ComplexTarget &target, bool conj, cudaStream_t st) {
    switch (dtype) {
        case DT_C64: return run_bound_complex<double2>(strided, X, ld, nvec, len, sft_out, target, conj, st);
        case DT_C32: return run_bound_complex<float2>(strided, X, ld, nvec, len, sft_out, target, conj, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_crt_complex(int dtype_C, bool split_weights, unsigned num_moduli, size_t m, size_t n, const uint8_t *C8u_re,
                               const uint8_t *C8u_im, size_t ldc8u, size_t sizeC, void *C, size_t ldc, const int16_t *sftA,
                               const int16_t *sftB, const void *alpha_p, const void *beta_p, bool device_scalars, cudaStream_t st) {
    if (m == 0 || n == 0) return cudaSuccess;
    if ((n + 3) / 4 > 65535) return cudaErrorInvalidValue;
    switch (dtype_C) {
        case DT_C64: return run_crt_cplx<double2>(split_weights, num_moduli, m, n, C8u_re, C8u_im, ldc8u, sizeC, C, ldc, sftA, sftB, alpha_p, beta_p, device_scalars, st);
        case DT_C32: return run_crt_cplx<float2>(false, num_moduli, m, n, C8u_re, C8u_im, ldc8u, sizeC, C, ldc, sftA, sftB, alpha_p, beta_p, device_scalars, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace oz
