/* CPU model of the DEVICE residue routes of mixed-gemmul8_b200/csrc/oz_residue.cuh (test infrastructure: lets the
 * exactness argument written there be checked against integer arithmetic without a GPU).  Same constants
 * (oz_tables.inc), same operation order; fma_rd is fma() under FE_DOWNWARD.
 *   residue_model_small_d / _f : floor quotient from the low word of fma_rd(a, rcp, 1.5 * 2^52 | 2^23), t = a - q m on the
 *                                low 32 bits, one fold "t > m/2 ? t - m : t"
 *   residue_model_big_d        : a = h 2^44 + l, v = h * (2^44 mod m) + l (exact, < 2^53), then the short route on v
 * Each returns the int8 the encoder would store.  */
#include <fenv.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#define OZ_TABLE(type, name, dims) static const type name dims
#include "../../mixed-gemmul8_b200/csrc/oz_tables.inc"
#undef OZ_TABLE

#pragma STDC FENV_ACCESS ON

static int32_t low_word_d(double a) { return (int32_t)(uint32_t)(uint64_t)(int64_t)a; }   /* cvt.rzi.s64.f64, low half */
static int32_t low_word_f(float a) { return (int32_t)a; }

static int32_t fold_down(int32_t t, int32_t half, int32_t m) { return t > half ? t - m : t; }

static double fma_rd(double a, double b, double c) {
    volatile double x = a, y = b, z = c, r;
    const int old = fegetround();
    fesetround(FE_DOWNWARD);
    r = fma(x, y, z);
    fesetround(old);
    return r;
}
static float fmaf_rd(float a, float b, float c) {
    volatile float x = a, y = b, z = c, r;
    const int old = fegetround();
    fesetround(FE_DOWNWARD);
    r = fmaf(x, y, z);
    fesetround(old);
    return r;
}
static int32_t lo32_of_double(double v) { uint64_t u; memcpy(&u, &v, 8); return (int32_t)(uint32_t)u; }

static int32_t small_d(double a, int32_t a_lo, unsigned j) {
    const int32_t m = OZ_MOD[j];
    if (j == 0) return a_lo;                                         /* m = 256: the low byte */
    const int32_t q = lo32_of_double(fma_rd(a, OZ_RCP64[j], 6755399441055744.0));
    const int32_t t = (int32_t)((uint32_t)q * (uint32_t)(-m) + (uint32_t)a_lo);
    return fold_down(t, m >> 1, m);
}
int residue_model_small_d(double a, unsigned j) { return (int8_t)small_d(a, low_word_d(a), j); }

int residue_model_small_f(float a, unsigned j) {
    const int32_t m = OZ_MOD[j], a_lo = low_word_f(a);
    if (j == 0) return (int8_t)a_lo;
    const float r = fmaf_rd(a, OZ_RCP32[j], 12582912.0f);
    uint32_t bits; memcpy(&bits, &r, 4);
    const int32_t q = (int32_t)bits - 0x4B400000;
    const int32_t t = (int32_t)((uint32_t)q * (uint32_t)(-m) + (uint32_t)a_lo);
    return (int8_t)fold_down(t, m >> 1, m);
}

int residue_model_big_d(double a, unsigned j) {
    const double h = trunc(a * 0x1p-44), l = fma(h, -17592186044416.0, a);
    const int32_t hlo = low_word_d(h), llo = low_word_d(l);
    if (j == 0) return (int8_t)llo;
    const double c = OZ_POW44[j];
    const double v = fma(h, c, l);                                       /* exact: |v| < 2^53 */
    const int32_t vlo = (int32_t)((uint32_t)hlo * (uint32_t)(int32_t)c + (uint32_t)llo);
    return (int8_t)small_d(v, vlo, j);
}
int residue_model_modulus(unsigned j) { return OZ_MOD[j]; }
