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
    // int -> fp through the exponent trick (exact; keeps I2F off the conversion pipe)
    c.neg_m  = 4503599627370496.0 - __hiloint2double(0x43300000, c.m);
    c.rcp    = OZ_RCP64[j];
    c.neg_mf = 8388608.0f - __int_as_float(0x4B000000 | c.m);
    c.rcpf   = OZ_RCP32[j];
    return c;
}
// The reference's operation sequence, instruction for instruction (any magnitude).
__device__ __forceinline__ int residue(double a, const ModConst &c) {
    float t = __double2float_rn(fma(rint(__dmul_rn(a, c.rcp)), c.neg_m, a));
    t       = __fmaf_rn(rintf(__fmul_rn(t, c.rcpf)), c.neg_mf, t);
    t       = __fmaf_rn(rintf(__fmul_rn(t, c.rcpf)), c.neg_mf, t);
    return __float2int_rz(t);
}
__device__ __forceinline__ int residue(float a, const ModConst &c) {
    float t = __fmaf_rn(rintf(__fmul_rn(a, c.rcpf)), c.neg_mf, a);
    t       = __fmaf_rn(rintf(__fmul_rn(t, c.rcpf)), c.neg_mf, t);
    t       = __fmaf_rn(rintf(__fmul_rn(t, c.rcpf)), c.neg_mf, t);
    t       = __fmaf_rn(rintf(__fmul_rn(t, c.rcpf)), c.neg_mf, t);
    return __float2int_rz(t);
}
// Same low byte for |a| below SmallLimit with ONE fp instruction per (element, modulus) and no
// conversion-pipe work (FRND / F2F / F2I issue at a quarter of the FP64 rate on sm_100 and bound
// the reference's sequence).  Why the result is identical:
//   * the reference ends on the symmetric residue of a mod m: its first remainder t = a - q*m is
//     an exact integer, and the float passes that follow subtract / add m until |t| <= m/2 (t/m is
//     a multiple of 1/m, so rintf can only tie for m = 256, where +-128 both wrap to int8 -128);
//   * so ANY quotient q with |a - q*m| <= 1.5 m followed by "fold once towards zero from either
//     side" lands on the same byte.  Here q = rint(a * rcp) comes out of the low word of
//     fma(a, rcp, 1.5*2^52) (|a * rcp| < 2^51; |a/m - q| <= 0.5 + 2^-53 |a/m| < 0.6), and
//     t = a - q*m is evaluated modulo 2^32 on the low words (|t| < 2^9, so nothing is lost).
template <typename R> struct SmallLimit;
template <> struct SmallLimit<double> { static constexpr double value = 0x1p57; };
template <> struct SmallLimit<float> { static constexpr float value = 0x1p24f; };
__device__ __forceinline__ int fold_once(int ti, int half, int m) {
    asm("{\n\t.reg .pred p, q;\n\t"
        "setp.gt.s32 p, %0, %1;\n\t"
        "@p sub.s32 %0, %0, %2;\n\t"
        "setp.lt.s32 q, %0, %3;\n\t"
        "@q add.s32 %0, %0, %2;\n\t}"
        : "+r"(ti)
        : "r"(half), "r"(m), "r"(-half));
    return ti;
}
// low 32 bits of the (exactly integer) value a, two's complement
__device__ __forceinline__ int low_word(double a) { return (int)__double2ll_rz(a); }
__device__ __forceinline__ int low_word(float a) { return __float2int_rz(a); }
__device__ __forceinline__ int residue_small(double a, int a_lo, const ModConst &c) {
    const int q = __double2loint(fma(a, c.rcp, 6755399441055744.0));  // 1.5 * 2^52
    return fold_once(q * c.neg_mi + a_lo, c.half, c.m);
}
__device__ __forceinline__ int residue_small(float a, int a_lo, const ModConst &c) {
    const int q = __float_as_int(__fmaf_rn(a, c.rcpf, 12582912.0f)) - 0x4B400000;  // 1.5 * 2^23
    return fold_once(q * c.neg_mi + a_lo, c.half, c.m);
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
};


}  // namespace
}  // namespace oz
