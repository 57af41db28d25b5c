// Scaling phase of the Ozaki-II emulation on B200: shift (power-of-two scale) selection per row of
// op(A) / column of op(B), and residue encoding of the scaled integers into int8 slices.
//
// What is computed follows the reference exactly (results are bit-identical):
//   fast-mode shift      GEMMul8/src/scaling.hpp:155-213 (reduction), :3373-3383 (compute_sft)
//   accurate-mode bound  GEMMul8/src/scaling.hpp:1897-1941, :2215-2260 (extract), :1504-1506 (compute_sft)
//   residue encode       GEMMul8/src/scaling.hpp:215-230 (mod_8i), :693-751, :1091-1148
// How it is computed is different: the reference runs one CTA per vector and gathers strided
// operands element by element; here strided operands are processed 32 vectors at a time with
// lane == vector (every warp load is one contiguous 256-byte segment), the encoder transposes
// through shared memory and emits 16-byte stores along K, and shifts / encode are separate kernels
// so the accurate mode reuses the encoder.
//
// Bit-exactness note.  The fast-mode bound uses a sum of squares accumulated with round-up FMAs,
// so its bits depend on the association order of the reference's reduction: thread t of a W-wide
// CTA (W = 128, or 512 for gemm<float>) accumulates elements t, t+W, t+2W, ... in order, warps are
// combined with a shfl_down tree and -- a quirk of the reference -- the per-warp value that is kept
// is the one that ends up in lane 1 (scaling.hpp:190-195), which omits lane 0's partial and counts
// lane 16's twice.  We reproduce that order with W "virtual threads" per vector.
#include "oz_common.cuh"
#include "oz_residue.cuh"

#include <cstdlib>

namespace oz {
namespace {

template <typename T> struct Real { using type = T; static constexpr bool cplx = false; };
template <> struct Real<float2> { using type = float; static constexpr bool cplx = true; };
template <> struct Real<double2> { using type = double; static constexpr bool cplx = true; };

__device__ __forceinline__ double fma_ru(double a, double b, double c) { return __fma_ru(a, b, c); }
__device__ __forceinline__ float fma_ru(float a, float b, float c) { return __fmaf_ru(a, b, c); }
__device__ __forceinline__ double add_ru(double a, double b) { return __dadd_ru(a, b); }
__device__ __forceinline__ float add_ru(float a, float b) { return __fadd_ru(a, b); }
__device__ __forceinline__ int ilogb_(double a) { return ilogb(a); }
__device__ __forceinline__ int ilogb_(float a) { return ilogbf(a); }

// the reference's warp trees (scaling.hpp:48-97): shfl_down by 16, 8, 4, 2, 1
template <typename R> __device__ __forceinline__ R tree_sum_ru(R v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = add_ru(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}
template <typename R> __device__ __forceinline__ R tree_max(R v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}

// reference: vecnorm::compute_sft, scaling.hpp:3373-3383.  A zero vector gets shift 0 (the
// reference leaves it undefined there; every product with it is exactly 0 either way).
__device__ __forceinline__ int fast_sft(double amax, double nrm2, float log2M) {
    if (amax == 0.0) return 0;
    const int e     = ilogb(nrm2);
    const float nf  = __double2float_ru(scalbn(nrm2, -e));
    const int k     = __float2int_rd(__fmaf_rd(-0.51f, __fadd_ru(__log2f(nf), (float)e), log2M));
    return min(__float2int_rd(log2M - 1.0f), k) - ilogb(amax);
}
__device__ __forceinline__ int fast_sft(float amax, float nrm2, float log2M) {
    if (amax == 0.0f) return 0;
    const int k = __float2int_rd(__fmaf_rd(-0.51f, __log2f(nrm2), log2M));
    return min(__float2int_rd(log2M - 1.0f), k) - ilogbf(amax);
}

template <typename T, typename R>
__device__ __forceinline__ void accumulate(const T &x, R &amax, R &acc) {
    if constexpr (Real<T>::cplx) {
        const R re = fabs(x.x), im = fabs(x.y);
        amax = fmax(amax, fmax(re, im));
        acc  = fma_ru(re, re, acc);
        acc  = fma_ru(im, im, acc);
    } else {
        const R a = fabs(x);
        amax = fmax(amax, a);
        acc  = fma_ru(a, a, acc);
    }
}

// ---------------------------------------------------------------------------------------------
// fast-mode shifts, contiguous vectors: one CTA of W threads per vector, literally the reference's
// thread ownership; loads are coalesced 4/8/16-byte per lane.
// ---------------------------------------------------------------------------------------------
template <typename T, int W>
__global__ void __launch_bounds__(W) fast_shift_contig_kernel(const T *__restrict__ X, size_t ld, size_t len,
                                                              float log2M, int16_t *__restrict__ out) {
    using R = typename Real<T>::type;
    __shared__ R s_max[32];
    __shared__ R s_sum[32];
    const T *__restrict__ p = X + (size_t)blockIdx.x * ld;
    R amax = 0, acc = 0;
    size_t i = threadIdx.x;
    // unrolled by 4 to keep several loads in flight; the FMA chain stays in element order
    for (; i + 3 * (size_t)W < len; i += 4 * (size_t)W) {
        const T x0 = p[i], x1 = p[i + W], x2 = p[i + 2 * (size_t)W], x3 = p[i + 3 * (size_t)W];
        accumulate<T, R>(x0, amax, acc);
        accumulate<T, R>(x1, amax, acc);
        accumulate<T, R>(x2, amax, acc);
        accumulate<T, R>(x3, amax, acc);
    }
    for (; i < len; i += W) accumulate<T, R>(p[i], amax, acc);

    amax = tree_max(amax);
    acc  = tree_sum_ru(acc);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_max[warp] = amax;
    if (lane == 1) s_sum[warp] = acc;  // lane 1, as the reference
    __syncthreads();
    if (warp == 0) {
        R a = (lane < W / 32) ? s_max[lane] : R(0);
        R s = (lane < W / 32) ? s_sum[lane] : R(0);
        a = tree_max(a);
        s = tree_sum_ru(s);
        if (lane == 0) out[blockIdx.x] = (int16_t)(-fast_sft(a, s, log2M));
    }
}

// ---------------------------------------------------------------------------------------------
// fast-mode shifts, strided vectors (rows of a column-major matrix): 32 vectors per CTA, lane ==
// vector (every warp load is one contiguous 32-element segment), W/32 warps.  Warp g plays the
// reference's warp g: its lanes' 32 virtual threads [32g, 32g+32) live in 32 registers of ONE
// thread, element i goes to accumulator (i mod W) in increasing i -- the reference's order -- and
// the reference's shuffle tree becomes register arithmetic (only the operands that reach lane 1
// are evaluated: lanes 1..31, lane 16 twice -- see the note at the top of this file).
// ---------------------------------------------------------------------------------------------
template <typename R> __device__ __forceinline__ R ref_tree_lane1(const R (&v)[32]) {
    R a[17], b[9];
#pragma unroll
    for (int i = 1; i <= 16; ++i) a[i] = add_ru(v[i], (i + 16 < 32) ? v[i + 16] : v[i]);   // o = 16, lanes 1..16
#pragma unroll
    for (int i = 1; i <= 8; ++i) b[i] = add_ru(a[i], a[i + 8]);                            // o = 8,  lanes 1..8
#pragma unroll
    for (int i = 1; i <= 4; ++i) a[i] = add_ru(b[i], b[i + 4]);                            // o = 4,  lanes 1..4
    b[1] = add_ru(a[1], a[3]);                                                              // o = 2,  lanes 1..2
    b[2] = add_ru(a[2], a[4]);
    return add_ru(b[1], b[2]);                                                              // o = 1,  lane 1
}
// second-level tree of the reference: NW warp values in lanes 0..NW-1, zeros above, lane 0 kept
template <typename R, int NW> __device__ __forceinline__ R ref_tree_lane0(const R *u, int stride) {
    R v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = (i < NW) ? u[i * stride] : R(0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int i = 0; i < o; ++i) v[i] = add_ru(v[i], v[i + o]);   // lanes >= o no longer feed lane 0
    }
    return v[0];
}

template <typename T, int W>
__global__ void __launch_bounds__(W) fast_shift_strided_kernel(const T *__restrict__ X, size_t ld, size_t nvec,
                                                               size_t len, float log2M, int16_t *__restrict__ out) {
    using R = typename Real<T>::type;
    constexpr int NW = W / 32;
    __shared__ R s_u[NW][33];
    __shared__ R s_amax[NW][33];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t vec  = (size_t)blockIdx.x * 32 + lane;
    const bool active = vec < nvec;
    const T *__restrict__ p = X + (active ? vec : 0);

    R acc[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) acc[q] = 0;
    R amax = 0;
    if (active) {
        size_t base = (size_t)warp * 32;
        for (; base + 32 <= len; base += W) {
            const T *__restrict__ pb = p + base * ld;
            constexpr int CH = sizeof(T) <= 8 ? 32 : 16;   // loads in flight per lane
#pragma unroll
            for (int q0 = 0; q0 < 32; q0 += CH) {
                T x[CH];
#pragma unroll
                for (int q = 0; q < CH; ++q) x[q] = pb[(size_t)(q0 + q) * ld];
#pragma unroll
                for (int q = 0; q < CH; ++q) accumulate<T, R>(x[q], amax, acc[q0 + q]);
            }
        }
        if (base < len) {
#pragma unroll
            for (int q = 0; q < 32; ++q)
                if (base + q < len) accumulate<T, R>(p[(base + q) * ld], amax, acc[q]);
        }
    }
    s_u[warp][lane]    = ref_tree_lane1(acc);
    s_amax[warp][lane] = amax;
    __syncthreads();
    if (warp == 0 && active) {
        R a = 0;
#pragma unroll
        for (int g = 0; g < NW; ++g) a = fmax(a, s_amax[g][lane]);
        const R s = ref_tree_lane0<R, NW>(&s_u[0][lane], 33);
        out[vec] = (int16_t)(-fast_sft(a, s, log2M));
    }
}

// The same with the 32 virtual threads of a virtual warp spread over 32 / VT real warps (VT accumulators per
// thread): W / VT warps per 32 vectors instead of W / 32, i.e. 4x (VT = 8) the loads in flight when there are few
// vectors -- 1024 ... 8192 rows give only 32 ... 256 CTAs, and the reference's order forbids splitting a vector's
// chains along k.  The partial sums meet in shared memory, where warp g evaluates virtual warp g's tree as above.
template <typename T, int W, int VT>
__global__ void __launch_bounds__(W / VT * 32) fast_shift_strided_split_kernel(const T *__restrict__ X, size_t ld, size_t nvec,
                                                                               size_t len, float log2M, int16_t *__restrict__ out) {
    using R = typename Real<T>::type;
    constexpr int NW = W / 32;            // virtual warps
    constexpr int NWARP = W / VT;         // real warps
    __shared__ R s_acc[W][32];            // [virtual thread][vector]
    __shared__ R s_amax[NWARP][32];
    __shared__ R s_u[NW][33];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t vec  = (size_t)blockIdx.x * 32 + lane;
    const bool active = vec < nvec;
    const T *__restrict__ p = X + (active ? vec : 0);

    R acc[VT];
#pragma unroll
    for (int q = 0; q < VT; ++q) acc[q] = 0;
    R amax = 0;
    if (active) {
        size_t base = (size_t)warp * VT;              // virtual threads [VT * warp, VT * warp + VT)
        for (; base + VT <= len; base += W) {
            const T *__restrict__ pb = p + base * ld;
            T x[VT];
#pragma unroll
            for (int q = 0; q < VT; ++q) x[q] = pb[(size_t)q * ld];
#pragma unroll
            for (int q = 0; q < VT; ++q) accumulate<T, R>(x[q], amax, acc[q]);
        }
        if (base < len) {
#pragma unroll
            for (int q = 0; q < VT; ++q)
                if (base + q < len) accumulate<T, R>(p[(base + q) * ld], amax, acc[q]);
        }
    }
#pragma unroll
    for (int q = 0; q < VT; ++q) s_acc[warp * VT + q][lane] = acc[q];
    s_amax[warp][lane] = amax;
    __syncthreads();
    if (warp < NW) {
        R v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = s_acc[warp * 32 + i][lane];
        s_u[warp][lane] = ref_tree_lane1(v);
    }
    __syncthreads();
    if (warp == 0 && active) {
        R a = 0;
#pragma unroll
        for (int g = 0; g < NWARP; ++g) a = fmax(a, s_amax[g][lane]);
        const R s = ref_tree_lane0<R, NW>(&s_u[0][lane], 33);
        out[vec] = (int16_t)(-fast_sft(a, s, log2M));
    }
}

// ---------------------------------------------------------------------------------------------
// encode, contiguous vectors.  One warp = 512 consecutive k of one vector: coalesced 16-byte loads (lane ==
// 16-byte chunk), a transpose through the warp's private shared-memory slab (XOR swizzle: conflict-free both
// ways, only __syncwarp), then each lane encodes 16 consecutive k and stores 16 bytes per modulus (a warp
// writes 4 full 128-byte lines).  16 elements per thread amortise the per-modulus constants and the store
// address over 16 residues (the 4-per-thread version spent a third of its issue slots there).
// grid = (ceil(ld8i / 512), ceil(nvec / 8)), 256 threads
// ---------------------------------------------------------------------------------------------
template <typename R, bool SPLIT>
__global__ void __launch_bounds__(256, SPLIT ? 2 : 3) encode_contig_kernel(const R *__restrict__ X, size_t ld, size_t nvec, size_t len,
                                                            const int16_t *__restrict__ sft_neg, unsigned num_moduli,
                                                            int8_t *__restrict__ out, size_t ld8i, size_t inc, bool aligned, bool ref_chain) {
    constexpr int EPC = 16 / sizeof(R);          // elements per 16-byte chunk
    constexpr int CPT = 16 / EPC;                // chunks per thread (16 elements)
    constexpr int CHUNKS = 32 * CPT;             // per warp tile of 512 elements
    __shared__ uint4 slab_all[8][CHUNKS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t vec = (size_t)blockIdx.y * 8 + warp;
    if (vec >= nvec) return;
    const size_t k0 = (size_t)blockIdx.x * 512;
    uint4 *__restrict__ slab = slab_all[warp];
    const R *__restrict__ p = X + vec * ld + k0;
    auto slot = [](int c) { return c ^ ((c >> 3) & (CPT - 1)); };
    if (aligned && k0 + 512 <= len) {
        uint4 x[CPT];
#pragma unroll
        for (int i = 0; i < CPT; ++i) x[i] = __ldg(reinterpret_cast<const uint4 *>(p) + i * 32 + lane);
#pragma unroll
        for (int i = 0; i < CPT; ++i) slab[slot(i * 32 + lane)] = x[i];
    } else {
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
            const int c = i * 32 + lane;
            R e[EPC];
#pragma unroll
            for (int q = 0; q < EPC; ++q) e[q] = (k0 + (size_t)c * EPC + q < len) ? p[(size_t)c * EPC + q] : R(0);
            slab[slot(c)] = *reinterpret_cast<const uint4 *>(e);
        }
    }
    __syncwarp();
    const size_t kb = k0 + 16 * lane;
    if (kb >= ld8i) return;
    R v[16];
#pragma unroll
    for (int j = 0; j < CPT; ++j) *reinterpret_cast<uint4 *>(&v[j * EPC]) = slab[slot(CPT * lane + j)];
    const Pow2<R> scale(-(int)sft_neg[vec]);     // one vector per warp: the rare two-factor case is warp-uniform
    if (scale.two) {
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = scale(v[e]);
    } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = scale.one(v[e]);
    }
    int8_t *__restrict__ o = out + vec * ld8i + kb;
    residues_of<16, SPLIT>(v, num_moduli, ref_chain, [&](unsigned j, const int (&q)[16]) {
        *reinterpret_cast<uint4 *>(o + (size_t)j * inc) =   // ld8i % 16 == 0 and kb % 16 == 0
            make_uint4(pack4(q[0], q[1], q[2], q[3]), pack4(q[4], q[5], q[6], q[7]), pack4(q[8], q[9], q[10], q[11]),
                       pack4(q[12], q[13], q[14], q[15]));
    });
}

// ---------------------------------------------------------------------------------------------
// encode, strided vectors: tile of 32 vectors x 128 k.  Load with lane == vector (256-byte
// contiguous segments), scale + truncate, park in shared memory with a rotation that makes both
// the column-wise writes and the row-wise reads conflict-free, then each thread encodes 16
// consecutive k of one vector and stores 16 bytes per modulus (a warp writes 4 full 128-byte lines).
// grid = (ceil(ld8i/128), ceil(nvec/32)), 256 threads
// ---------------------------------------------------------------------------------------------
template <typename R, bool SPLIT>
__global__ void __launch_bounds__(256, SPLIT ? 2 : 4) encode_strided_kernel(const R *__restrict__ X, size_t ld, size_t nvec, size_t len,
                                                             const int16_t *__restrict__ sft_neg, unsigned num_moduli,
                                                             int8_t *__restrict__ out, size_t ld8i, size_t inc, bool ref_chain) {
    __shared__ R tile[128 * 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t v0 = (size_t)blockIdx.y * 32;
    const size_t k0 = (size_t)blockIdx.x * 128;
    if (v0 + 32 <= nvec && k0 + 128 <= len) {
        // interior tile: no bounds checks, one 64-bit pointer stepped by 8 rows, all 16 loads in flight
        const Pow2<R> scale(-(int)sft_neg[v0 + lane]);
        const R *__restrict__ p = X + v0 + lane + (k0 + warp) * ld;
        const size_t step = 8 * ld;
        R x[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { x[i] = *p; p += step; }
        R *__restrict__ t = tile + warp * 32;
        if (__any_sync(0xffffffffu, scale.two)) {   // shifts beyond +-1022: two exact factors
#pragma unroll
            for (int i = 0; i < 16; ++i) t[i * 8 * 32 + ((lane + 4 * (i >> 1)) & 31)] = scale(x[i]);
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)    // kk = warp + 8 i, so kk >> 4 == i >> 1
                t[i * 8 * 32 + ((lane + 4 * (i >> 1)) & 31)] = scale.one(x[i]);
        }
    } else {
        const size_t vec = v0 + lane;
        const bool active = vec < nvec;
        const Pow2<R> scale(active ? -(int)sft_neg[vec] : 0);
        const R *__restrict__ p = X + (active ? vec : 0);
#pragma unroll 8
        for (int kk = warp; kk < 128; kk += 8) {
            const size_t k = k0 + kk;
            R x = (active && k < len) ? p[k * ld] : R(0);
            tile[kk * 32 + ((lane + 4 * (kk >> 4)) & 31)] = scale(x);
        }
    }
    __syncthreads();
    const int r = threadIdx.x >> 3, g = threadIdx.x & 7;  // vector in tile, 16-wide k group
    const size_t vec = v0 + r;
    const size_t kb  = k0 + 16 * g;
    if (vec >= nvec || kb >= ld8i) return;
    R v[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = tile[(16 * g + e) * 32 + ((r + 4 * g) & 31)];
    int8_t *__restrict__ o = out + vec * ld8i + kb;
    residues_of<16, SPLIT>(v, num_moduli, ref_chain, [&](unsigned j, const int (&q)[16]) {
        *reinterpret_cast<uint4 *>(o + (size_t)j * inc) =   // ld8i % 16 == 0 and kb % 16 == 0
            make_uint4(pack4(q[0], q[1], q[2], q[3]), pack4(q[4], q[5], q[6], q[7]), pack4(q[8], q[9], q[10], q[11]),
                       pack4(q[12], q[13], q[14], q[15]));
    });
}

// ---------------------------------------------------------------------------------------------
// accurate mode, step 1: amax -> sft0 = 5 - ilogb(amax); bound slice = ceil(|x| * 2^sft0) (<= 64)
// reference: extract_A8i_kernel / extract_B8i_kernel, scaling.hpp:1897-1941, :2215-2260
// Two kernels: amax per vector (exact, order-free), then the element-wise extraction.
// ---------------------------------------------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(256) amax_contig_kernel(const R *__restrict__ X, size_t ld, size_t len,
                                                          int16_t *__restrict__ sft0) {
    __shared__ R s_max[8];
    const R *__restrict__ p = X + (size_t)blockIdx.x * ld;
    R amax = 0;
    for (size_t i = threadIdx.x; i < len; i += 256) amax = fmax(amax, fabs(p[i]));
    amax = tree_max(amax);
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = amax;
    __syncthreads();
    if (threadIdx.x < 32) {
        R a = (threadIdx.x < 8) ? s_max[threadIdx.x] : R(0);
        a   = tree_max(a);
        if (threadIdx.x == 0) sft0[blockIdx.x] = (a == R(0)) ? (int16_t)0 : (int16_t)(5 - ilogb_(a));
    }
}
template <typename R>
__global__ void __launch_bounds__(512) amax_strided_kernel(const R *__restrict__ X, size_t ld, size_t nvec, size_t len,
                                                           int16_t *__restrict__ sft0) {
    __shared__ R s_max[16 * 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t vec = (size_t)blockIdx.x * 32 + lane;
    R amax = 0;
    if (vec < nvec) {
        const R *__restrict__ p = X + vec;
        for (size_t i = warp; i < len; i += 16) amax = fmax(amax, fabs(p[i * ld]));
    }
    s_max[warp * 32 + lane] = amax;
    __syncthreads();
    if (warp == 0 && vec < nvec) {
        R a = 0;
#pragma unroll
        for (int w = 0; w < 16; ++w) a = fmax(a, s_max[w * 32 + lane]);
        sft0[vec] = (a == R(0)) ? (int16_t)0 : (int16_t)(5 - ilogb_(a));
    }
}
__device__ __forceinline__ int bound_int(double x, int sft) { return __double2int_ru(scalbn(fabs(x), sft)); }
__device__ __forceinline__ int bound_int(float x, int sft) { return __float2int_ru(scalbnf(fabsf(x), sft)); }

template <typename R>
__global__ void __launch_bounds__(256) bound_contig_kernel(const R *__restrict__ X, size_t ld, size_t len,
                                                           const int16_t *__restrict__ sft0, int8_t *__restrict__ out,
                                                           size_t ld8i) {
    const size_t vec = blockIdx.y;
    const size_t i0  = ((size_t)blockIdx.x * 256 + threadIdx.x) * 4;
    if (i0 >= ld8i) return;
    const int sft = sft0[vec];
    const R *__restrict__ p = X + vec * ld;
    int b[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) b[e] = (i0 + e < len) ? bound_int(p[i0 + e], sft) : 0;
    *reinterpret_cast<uint32_t *>(out + vec * ld8i + i0) = pack4(b[0], b[1], b[2], b[3]);
}
template <typename R>
__global__ void __launch_bounds__(256) bound_strided_kernel(const R *__restrict__ X, size_t ld, size_t nvec, size_t len,
                                                            const int16_t *__restrict__ sft0, int8_t *__restrict__ out,
                                                            size_t ld8i) {
    __shared__ int8_t tile[32][128 + 16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t v0 = (size_t)blockIdx.y * 32, k0 = (size_t)blockIdx.x * 128;
    {
        const size_t vec = v0 + lane;
        const bool active = vec < nvec;
        const int sft = active ? (int)sft0[vec] : 0;
        const R *__restrict__ p = X + (active ? vec : 0);
        for (int kk = warp; kk < 128; kk += 8) {
            const size_t k = k0 + kk;
            tile[lane][kk] = (int8_t)((active && k < len) ? bound_int(p[k * ld], sft) : 0);
        }
    }
    __syncthreads();
    const int r = threadIdx.x >> 3, g = threadIdx.x & 7;
    const size_t vec = v0 + r, kb = k0 + 16 * g;
    if (vec >= nvec || kb >= ld8i) return;
    *reinterpret_cast<uint4 *>(out + vec * ld8i + kb) = *reinterpret_cast<const uint4 *>(&tile[r][16 * g]);
}

// accurate mode, step 3 (reference: int8tc::compute_sft, scaling.hpp:1504-1506)
__global__ void accurate_shift_kernel(size_t nvec, const int32_t *__restrict__ cmax, float log2M,
                                      int16_t *__restrict__ sft, size_t fold) {
    const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvec) return;
    const int s0 = sft[v];
    const int cm = fold ? max(cmax[v], cmax[v + fold]) : cmax[v];
    // cmax == 0 (a zero vector, or every partner vector is zero): any shift gives exact zeros; keep
    // the 6-bit one (the reference evaluates log2(0) here and stores an unspecified shift)
    const int s  = (cm <= 0) ? s0 : s0 + __float2int_rd(__fmaf_rd(-0.51f, __log2f(__int2float_rn(cm)), log2M));
    sft[v]       = (int16_t)(-s);
}

template <typename T, int W>
cudaError_t run_fast_shifts(bool strided, const void *X, size_t ld, size_t nvec, size_t len, float log2M,
                            int16_t *out, cudaStream_t st) {
    using R = typename Real<T>::type;
    if (nvec == 0) return cudaSuccess;
    if (strided) {
        static const int split_below = [] { const char *e = getenv("OZ_SHIFT_SPLIT_BELOW"); return e ? atoi(e) : 32768; }();
        if constexpr (W == 128 && sizeof(R) <= 8 && !Real<T>::cplx) {
            if (nvec < (size_t)split_below) {
                fast_shift_strided_split_kernel<T, W, 8><<<(unsigned)((nvec + 31) / 32), W / 8 * 32, 0, st>>>(static_cast<const T *>(X), ld, nvec,
                                                                                                            len, log2M, out);
                count_launch();
                return cudaGetLastError();
            }
        }
        fast_shift_strided_kernel<T, W><<<(unsigned)((nvec + 31) / 32), W, 0, st>>>(static_cast<const T *>(X), ld, nvec, len, log2M, out);
    } else {
        fast_shift_contig_kernel<T, W><<<(unsigned)nvec, W, 0, st>>>(static_cast<const T *>(X), ld, len, log2M, out);
    }
    count_launch();
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_fast_shifts(int dtype, bool strided, const void *X, size_t ld, size_t nvec, size_t len,
                               int ref_width, float log2M, int16_t *sft_out, cudaStream_t st) {
    const bool w512 = ref_width == 512;
    switch (dtype) {
        case DT_F64: return run_fast_shifts<double, 128>(strided, X, ld, nvec, len, log2M, sft_out, st);  // 512 only occurs for gemm<float>
        case DT_F32: return w512 ? run_fast_shifts<float, 512>(strided, X, ld, nvec, len, log2M, sft_out, st)
                                 : run_fast_shifts<float, 128>(strided, X, ld, nvec, len, log2M, sft_out, st);
        case DT_C64: return run_fast_shifts<double2, 128>(strided, X, ld, nvec, len, log2M, sft_out, st);
        case DT_C32: return run_fast_shifts<float2, 128>(strided, X, ld, nvec, len, log2M, sft_out, st);
    }
    return cudaErrorInvalidValue;
}

template <typename R>
static cudaError_t run_encode(bool strided, const void *X, size_t ld, size_t nvec, size_t len, const int16_t *sft_neg,
                              unsigned num_moduli, int8_t *out, size_t ld8i, size_t inc, cudaStream_t st) {
    const bool ref_chain = encode_reference_chain();
    if (nvec == 0) return cudaSuccess;
    if (strided) {
        dim3 grid((unsigned)((ld8i + 127) / 128), (unsigned)((nvec + 31) / 32));
        auto kern = num_moduli >= 16 ? encode_strided_kernel<R, true> : encode_strided_kernel<R, false>;
        kern<<<grid, 256, 0, st>>>(static_cast<const R *>(X), ld, nvec, len, sft_neg, num_moduli, out, ld8i, inc, ref_chain);
        count_launch();
    } else {
        // blockIdx.y is limited to 65535: walk the vectors in slabs of 8 x 65535
        const bool aligned = (reinterpret_cast<uintptr_t>(X) % 16 == 0) && (ld * sizeof(R)) % 16 == 0;
        constexpr size_t SLAB = 8 * (size_t)65535;
        for (size_t v0 = 0; v0 < nvec; v0 += SLAB) {
            const size_t nv = nvec - v0 < SLAB ? nvec - v0 : SLAB;
            dim3 grid((unsigned)((ld8i + 511) / 512), (unsigned)((nv + 7) / 8));
            auto kern = num_moduli >= 16 ? encode_contig_kernel<R, true> : encode_contig_kernel<R, false>;
            kern<<<grid, 256, 0, st>>>(static_cast<const R *>(X) + v0 * ld, ld, nv, len, sft_neg + v0, num_moduli, out + v0 * ld8i,
                                       ld8i, inc, aligned, ref_chain);
            count_launch();
        }
    }
    return cudaGetLastError();
}

cudaError_t launch_encode(int dtype, bool strided, const void *X, size_t ld, size_t nvec, size_t len,
                          const int16_t *sft_neg, unsigned num_moduli, int8_t *out, size_t ld8i, size_t inc,
                          cudaStream_t st) {
    if (strided && (nvec + 31) / 32 > 65535) return cudaErrorInvalidValue;
    switch (dtype) {
        case DT_F64: return run_encode<double>(strided, X, ld, nvec, len, sft_neg, num_moduli, out, ld8i, inc, st);
        case DT_F32: return run_encode<float>(strided, X, ld, nvec, len, sft_neg, num_moduli, out, ld8i, inc, st);
    }
    return cudaErrorInvalidValue;
}

template <typename R>
static cudaError_t run_bound(bool strided, const void *X, size_t ld, size_t nvec, size_t len, int8_t *out, size_t ld8i,
                             int16_t *sft0, cudaStream_t st) {
    if (nvec == 0) return cudaSuccess;
    const R *x = static_cast<const R *>(X);
    if (strided) {
        amax_strided_kernel<R><<<(unsigned)((nvec + 31) / 32), 512, 0, st>>>(x, ld, nvec, len, sft0);
        dim3 grid((unsigned)((ld8i + 127) / 128), (unsigned)((nvec + 31) / 32));
        bound_strided_kernel<R><<<grid, 256, 0, st>>>(x, ld, nvec, len, sft0, out, ld8i);
        count_launch(2);
    } else {
        amax_contig_kernel<R><<<(unsigned)nvec, 256, 0, st>>>(x, ld, len, sft0);
        count_launch();
        for (size_t v0 = 0; v0 < nvec; v0 += 65535) {
            const size_t nv = nvec - v0 < 65535 ? nvec - v0 : 65535;
            dim3 grid((unsigned)((ld8i / 4 + 255) / 256), (unsigned)nv);
            bound_contig_kernel<R><<<grid, 256, 0, st>>>(x + v0 * ld, ld, len, sft0 + v0, out + v0 * ld8i, ld8i);
            count_launch();
        }
    }
    return cudaGetLastError();
}

cudaError_t launch_bound_extract(int dtype, bool strided, const void *X, size_t ld, size_t nvec, size_t len,
                                 int8_t *out8i, size_t ld8i, int16_t *sft_out, cudaStream_t st) {
    switch (dtype) {
        case DT_F64: return run_bound<double>(strided, X, ld, nvec, len, out8i, ld8i, sft_out, st);
        case DT_F32: return run_bound<float>(strided, X, ld, nvec, len, out8i, ld8i, sft_out, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_accurate_shifts(size_t nvec, const int32_t *cmax, float log2M, int16_t *sft_inout,
                                   cudaStream_t st, size_t fold) {
    if (nvec == 0) return cudaSuccess;
    accurate_shift_kernel<<<(unsigned)((nvec + 255) / 256), 256, 0, st>>>(nvec, cmax, log2M, sft_inout, fold);
    count_launch();
    return cudaGetLastError();
}

}  // namespace oz
