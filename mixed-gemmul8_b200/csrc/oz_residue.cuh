// Residue arithmetic shared by the real (oz_scale.cu) and complex (oz_complex.cu) encoders.
#pragma once
#include "oz_common.cuh"

namespace oz {
namespace {

using dev_tab::OZ_MOD;
using dev_tab::OZ_RCP32;
using dev_tab::OZ_RCP64;

// ---------------------------------------------------------------------------------------------
// residue of an integer-valued fp number modulo m_j, symmetric representative, as int8
// reference: mod_8i, scaling.hpp:215-230 (cast wraps: cvt.rzi.s32.f32 + byte store)
// ---------------------------------------------------------------------------------------------
struct ModConst {
    double neg_m, rcp;
    float neg_mf, rcpf;
    int m, half, neg_mi;
};
__device__ __forceinline__ ModConst load_mod(unsigned j) {
    ModConst c;
    c.m      = OZ_MOD[j];
    c.half   = c.m >> 1;
    c.neg_mi = -c.m;
    asm volatile("" : "+r"(c.neg_mi));   // keep -m in a register: otherwise every q * (-m) becomes a negate + IMAD
    // int -> fp through the exponent trick (exact; keeps I2F off the conversion pipe)
    c.neg_m  = 4503599627370496.0 - __hiloint2double(0x43300000, c.m);
    c.rcp    = OZ_RCP64[j];
    c.neg_mf = 8388608.0f - __int_as_float(0x4B000000 | c.m);
    c.rcpf   = OZ_RCP32[j];
    return c;
}
// The reference's operation sequence (scaling.hpp:215-230).  It ends on the exact symmetric residue
// for every magnitude the scaling can produce (rint(a * rcp) is within a few units of a / m, every
// fma in the chain is exact and the float passes fold the rest; checked against integer arithmetic
// up to 2^79 by a CPU property test under tests/), so the encoders may take ANY exact route to the same byte.
// Kept for the on-device cross-check (GEMMUL8_FLAG_ENCODE_REFERENCE).
__device__ __forceinline__ int residue_reference(double a, const ModConst &c) {
    float t = __double2float_rn(fma(rint(__dmul_rn(a, c.rcp)), c.neg_m, a));
    t       = __fmaf_rn(rintf(__fmul_rn(t, c.rcpf)), c.neg_mf, t);
    t       = __fmaf_rn(rintf(__fmul_rn(t, c.rcpf)), c.neg_mf, t);
    return __float2int_rz(t);
}
__device__ __forceinline__ int residue_reference(float a, const ModConst &c) {
    float t = __fmaf_rn(rintf(__fmul_rn(a, c.rcpf)), c.neg_mf, a);
    t       = __fmaf_rn(rintf(__fmul_rn(t, c.rcpf)), c.neg_mf, t);
    t       = __fmaf_rn(rintf(__fmul_rn(t, c.rcpf)), c.neg_mf, t);
    t       = __fmaf_rn(rintf(__fmul_rn(t, c.rcpf)), c.neg_mf, t);
    return __float2int_rz(t);
}
// The routes used instead: one fp instruction per (element, modulus) and no conversion-pipe work
// (FRND / F2F / F2I issue at a quarter of the FP64 rate on sm_100 and bound the sequence above).
//   * q = floor(a * rcp) comes out of the low word of fma_rd(a, rcp, 1.5*2^52) (needs |a * rcp| < 2^51).  rcp is
//     1/m rounded, so the product is off by e, |e| <= 2^-53 |a/m| < 0.1 for |a| < 2^57 (fp32: 2^-24 |a/m| < 0.01
//     for |a| < 2^24), and q is floor(a/m) or a neighbour:
//         q = floor(a/m)     : t = a - q m = r in [0, m)
//         q = floor(a/m) - 1 : only if r/m + e < 0,  i.e. r < 0.1 m : t = r + m
//         q = floor(a/m) + 1 : only if r/m + e >= 1, i.e. r > 0.9 m : t = r - m, already the symmetric residue
//     so ONE fold, "t > m/2 ? t - m : t", lands on the symmetric residue in every case (m = 256: r = 128 stays
//     +128, the same int8 as the reference's -128).  t is evaluated modulo 2^32 on the low words (|t| < 2^10).
//     Valid for |a| < 2^57 (fp64) / 2^24 (fp32);
//   * m = 256 (modulus 0) needs no arithmetic at all: the int8 is the low byte of a;
//   * larger values (more than 15 moduli; |a| < 2^89) are split exactly as a = h * 2^44 + l, |h| < 2^45, |l| < 2^44, and
//     folded BEFORE the reduction: v = h * c + l with c = 2^44 mod m (symmetric, |c| <= 128) is congruent to a, an exact
//     integer below 2^53 (one DFMA), and takes the short route; its low word is h_lo * c + l_lo in 32-bit arithmetic.
//     Two FP64 instructions and two IMADs per (element, modulus) -- the first version reduced h and l separately and
//     joined the residues (three reductions, ~18 instructions), which made the encoder compute-bound from 17 moduli on
//     (scaling phase 15 ms instead of 3.6 at 16384^2 x 16384).
template <typename R> struct SmallLimit;
template <> struct SmallLimit<double> { static constexpr double value = 0x1p57; };
template <> struct SmallLimit<float> { static constexpr float value = 0x1p24f; };
__device__ __forceinline__ int fold_down(int ti, int half, int m) {   // (-m/2, m + m/2) -> symmetric
    asm("{\n\t.reg .pred p;\n\t"
        "setp.gt.s32 p, %0, %1;\n\t"
        "@p sub.s32 %0, %0, %2;\n\t}"
        : "+r"(ti)
        : "r"(half), "r"(m));
    return ti;
}
// low 32 bits of the (exactly integer) value a, two's complement
__device__ __forceinline__ int low_word(double a) { return (int)__double2ll_rz(a); }
__device__ __forceinline__ int low_word(float a) { return __float2int_rz(a); }
__device__ __forceinline__ int residue_small(double a, int a_lo, const ModConst &c) {
    const int q = __double2loint(__fma_rd(a, c.rcp, 6755399441055744.0));  // 1.5 * 2^52
    return fold_down(q * c.neg_mi + a_lo, c.half, c.m);
}
__device__ __forceinline__ int residue_small(float a, int a_lo, const ModConst &c) {
    const int q = __float_as_int(__fmaf_rd(a, c.rcpf, 12582912.0f)) - 0x4B400000;  // 1.5 * 2^23
    return fold_down(q * c.neg_mi + a_lo, c.half, c.m);
}
// |a| < 2^89 through the halves h = trunc(a / 2^44), l = a - h * 2^44: v = h * (2^44 mod m) + l is exact and congruent to a
__device__ __forceinline__ int residue_big(double h, int h_lo, double l, int l_lo, double pow44, int pow44_i, const ModConst &c) {
    const double v = fma(h, pow44, l);                   // |v| <= 2^45 * 128 + 2^44 < 2^53: exact
    return residue_small(v, h_lo * pow44_i + l_lo, c);   // low word of v modulo 2^32
}

// Residues of G integer-valued elements for every modulus; `store(j, r)` receives the G residues of
// modulus j (the low byte of each int is the int8 to write).  Per thread, one of the exact routes.
// SPLIT = false: values beyond the short route take the reference's chain (rare below 16 moduli, and it
// keeps the kernel at 64 registers); SPLIT = true (launched for 16+ moduli): they take the split route.
template <int G, bool SPLIT, typename Store>
__device__ __forceinline__ void residues_double(const double (&v)[G], unsigned num_moduli, bool reference_chain, Store &&store) {
    unsigned top = 0;   // |v| < 2^57 on the high words (integer compare: keeps the test off the FP64 pipe)
#pragma unroll
    for (int e = 0; e < G; ++e) top = max(top, (unsigned)__double2hiint(v[e]) & 0x7fffffffu);
    // (SPLIT kernels: one decision per warp -- the long route is valid for small values too, and a warp whose lanes disagree
    //  would execute both; with 17 moduli about half of the threads of a warp hold a value beyond 2^57)
    const bool small = SPLIT ? __all_sync(__activemask(), top < 0x43800000u) != 0 : top < 0x43800000u;
    if (reference_chain || (!SPLIT && !small)) {
        for (unsigned j = 0; j < num_moduli; ++j) {
            const ModConst c = load_mod(j);
            int r[G];
#pragma unroll
            for (int e = 0; e < G; ++e) r[e] = residue_reference(v[e], c);
            store(j, r);
        }
    } else if (small) {
        int lo[G];
#pragma unroll
        for (int e = 0; e < G; ++e) {
            lo[e] = low_word(v[e]);
            asm volatile("" : "+r"(lo[e]));   // computed once: do not rematerialise the F2I inside the modulus loop
        }
        store(0u, lo);   // m_0 = 256: the low byte
        for (unsigned j = 1; j < num_moduli; ++j) {
            const ModConst c = load_mod(j);
            int r[G];
#pragma unroll
            for (int e = 0; e < G; ++e) r[e] = residue_small(v[e], lo[e], c);
            store(j, r);
        }
    } else if constexpr (SPLIT) {
        double h[G], l[G];
        int hlo[G], llo[G];
#pragma unroll
        for (int e = 0; e < G; ++e) {
            h[e]   = trunc(v[e] * 0x1p-44);
            l[e]   = fma(h[e], -17592186044416.0, v[e]);
            hlo[e] = low_word(h[e]);
            llo[e] = low_word(l[e]);
            asm volatile("" : "+r"(hlo[e]), "+r"(llo[e]));
        }
        store(0u, llo);  // 2^44 = 0 mod 256
        for (unsigned j = 1; j < num_moduli; ++j) {
            const ModConst c = load_mod(j);
            const double p44 = dev_tab::OZ_POW44[j];
            const int p44i   = (int)p44;
            int r[G];
#pragma unroll
            for (int e = 0; e < G; ++e) r[e] = residue_big(h[e], hlo[e], l[e], llo[e], p44, p44i, c);
            store(j, r);
        }
    }
}
template <int G, bool SPLIT, typename Store>
__device__ __forceinline__ void residues_of(const double (&v)[G], unsigned num_moduli, bool reference_chain, Store &&store) {
    residues_double<G, SPLIT>(v, num_moduli, reference_chain, store);
}
template <int G, bool SPLIT, typename Store>
__device__ __forceinline__ void residues_of(const float (&v)[G], unsigned num_moduli, bool reference_chain, Store &&store) {
    unsigned top = 0;   // |v| < 2^24
#pragma unroll
    for (int e = 0; e < G; ++e) top = max(top, (unsigned)__float_as_int(v[e]) & 0x7fffffffu);
    const bool small = top < 0x4B800000u;
    if (reference_chain) {
        for (unsigned j = 0; j < num_moduli; ++j) {
            const ModConst c = load_mod(j);
            int r[G];
#pragma unroll
            for (int e = 0; e < G; ++e) r[e] = residue_reference(v[e], c);
            store(j, r);
        }
    } else if (small) {
        int lo[G];
#pragma unroll
        for (int e = 0; e < G; ++e) {
            lo[e] = low_word(v[e]);
            asm volatile("" : "+r"(lo[e]));
        }
        store(0u, lo);
        for (unsigned j = 1; j < num_moduli; ++j) {
            const ModConst c = load_mod(j);
            int r[G];
#pragma unroll
            for (int e = 0; e < G; ++e) r[e] = residue_small(v[e], lo[e], c);
            store(j, r);
        }
    } else {   // an fp32 value beyond 2^24 is still an exact integer: widen and take the fp64 routes
        double d[G];
#pragma unroll
        for (int e = 0; e < G; ++e) d[e] = (double)v[e];
        residues_double<G, SPLIT>(d, num_moduli, false, store);
    }
}

__device__ __forceinline__ uint32_t pack4(int a, int b, int c, int d) {  // low bytes, 3 PRMT
    return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}
// trunc(x * 2^sft): scalbn by an exact power-of-two product.  One factor when 2^sft is a normal
// number, two otherwise (tiny operands); a result too small to be normal truncates to 0 either way.
template <typename R> struct Pow2 {
    R f1, f2;
    bool two;
    __device__ __forceinline__ explicit Pow2(int sft) {
        constexpr int lim = sizeof(R) == 8 ? 1022 : 126;
        two = sft > lim || sft < -lim;
        const int s1 = two ? sft / 2 : sft, s2 = sft - s1;
        if constexpr (sizeof(R) == 8) {
            f1 = __hiloint2double((1023 + s1) << 20, 0);
            f2 = __hiloint2double((1023 + s2) << 20, 0);
        } else {
            f1 = __int_as_float((127 + s1) << 23);
            f2 = __int_as_float((127 + s2) << 23);
        }
    }
    __device__ __forceinline__ R operator()(R x) const {
        R y = x * f1;
        if (two) y *= f2;
        if constexpr (sizeof(R) == 8) return trunc(y); else return truncf(y);
    }
    // the same when `two` is known to be false (callers branch once per warp: keeps the second multiply and its
    // selects out of the common path)
    __device__ __forceinline__ R one(R x) const {
        if constexpr (sizeof(R) == 8) return trunc(x * f1); else return truncf(x * f1);
    }
};


}  // namespace
}  // namespace oz
