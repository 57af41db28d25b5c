#!/usr/bin/env python3
"""Generate the Ozaki-scheme-II constant tables used by the CUDA kernels and the CPU oracle.

Everything here is derived from first principles with exact integer arithmetic:

  moduli m_j        : the 20 pairwise-coprime moduli <= 256 (fixed by the number format)
  1/m_j             : fp64 / fp32 reciprocals, round-to-nearest
  M = prod m_j      : as an unevaluated (hi, lo) double pair, hi = RN(M), lo = RN(M - hi)
  1/M               : RN
  log2M budgets     : float, rounded down: log2(M-1)/2 - 0.5 (accurate) and - 1.5 (fast)
  CRT weights       : w_j = (M/m_j) * ((M/m_j)^-1 mod m_j)
      single  (N<=7 or fp32 output) : one double per weight
      split   (N>=8, fp64 output)   : hi = w_j truncated to a multiple of 2^g with
                                      g = bitlen(M) - 44 + ceil(log2 N), so that
                                      sum_j hi_j * r_j (r_j <= 255) is exact in fp64;
                                      lo = w_j - hi

The reference (GEMMul8/src/table.hpp:27-826) ships the same quantities as decimal literals.  A
few of its entries (lo parts wider than 53 bits, single weights for N >= 8) were produced by an
extended-precision tool and differ from the correctly rounded value by a few ulps.  Because the
final fp64 C of the reference depends on those exact doubles, `tools/ref_ulp_deltas.json` records
the (tiny, integer) ulp offsets so that this generator reproduces the reference's doubles
bit-for-bit.  `--check-ref` re-derives the offsets from /root/reference (only available in the
build container) and verifies every other entry against the reference's literals.

Outputs: mixed-gemmul8_b200/csrc/oz_tables.inc (C/C++ initialisers shared by CUDA and the oracle).
"""
import argparse
import json
import math
import os
import re
import struct
from fractions import Fraction

MODULI = [256, 255, 253, 251, 247, 241, 239, 233, 229, 227, 223, 217, 211, 199, 197, 193, 191, 181, 179, 173]
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DELTAS = os.path.join(HERE, "ref_ulp_deltas.json")
OUT = os.path.join(ROOT, "mixed-gemmul8_b200", "csrc", "oz_tables.inc")


def f32(x):
    """Round a Python float / Fraction to the nearest binary32 (returned as Python float)."""
    return struct.unpack("f", struct.pack("f", float(x)))[0]


def rn_fraction(fr, bits=53):
    """Correctly rounded (nearest-even) conversion of a positive Fraction to a float with `bits` precision."""
    if fr == 0:
        return 0.0
    sign = -1 if fr < 0 else 1
    fr = abs(fr)
    e = fr.numerator.bit_length() - fr.denominator.bit_length()
    # normalise so that 2^(bits-1) <= q < 2^bits
    shift = bits - 1 - e
    scaled = fr * (Fraction(2) ** shift)
    if scaled < 2 ** (bits - 1):
        shift += 1
        scaled *= 2
    elif scaled >= 2 ** bits:
        shift -= 1
        scaled /= 2
    q, r = divmod(scaled.numerator, scaled.denominator)
    twice = 2 * r
    if twice > scaled.denominator or (twice == scaled.denominator and (q & 1)):
        q += 1
    return sign * math.ldexp(q, -shift)


def rd_float32(fr_lo, fr_hi):
    """Largest binary32 <= the real number bracketed by [fr_lo, fr_hi] (must agree for both ends)."""
    def rd(fr):
        x = f32(rn_fraction(fr, 24))
        if Fraction(x) > fr:
            x = f32(math.nextafter(x, -math.inf)) if False else struct.unpack("f", struct.pack("I", struct.unpack("I", struct.pack("f", x))[0] - 1))[0]
        return x
    a, b = rd(fr_lo), rd(fr_hi)
    assert a == b, "log2 bracket straddles a float boundary; widen precision"
    return a


def dec9_float32(fr_lo, fr_hi):
    """The reference's accurate-mode budget literals are the exact value cut to 9 significant decimal
    digits (toward zero) and then read by the compiler as the nearest binary32."""
    def cut(fr):
        digits = len(str(int(fr)))                # integer digits (values are in [1, 100))
        scale = 10 ** (9 - digits)
        return Fraction(int(fr * scale), scale)
    a, b = cut(fr_lo), cut(fr_hi)
    assert a == b, "log2 bracket straddles a decimal boundary; widen precision"
    return f32(rn_fraction(a, 24))


def log2_bracket(n, prec=200):
    """Rational lower/upper bounds of log2(n) for a positive integer n, error < 2^-prec."""
    e = n.bit_length() - 1
    # log2(n) = e + log2(n / 2^e), mantissa x in [1,2): square-and-compare, prec bits
    x = Fraction(n, 1 << e)
    lo = Fraction(e)
    step = Fraction(1, 2)
    for _ in range(prec):
        x = x * x
        # keep the fraction small: truncate to ~4*prec bits (floor keeps a lower bound)
        if x.denominator.bit_length() > 4 * prec:
            sh = x.denominator.bit_length() - 4 * prec
            x = Fraction(x.numerator >> sh, x.denominator >> sh)
        if x >= 2:
            x /= 2
            lo += step
        step /= 2
    return lo, lo + 4 * step * prec  # generous upper bound (truncation only lowers x)


def ulp_add(x, n):
    """x displaced by n units in the last place (binary64)."""
    for _ in range(abs(n)):
        x = math.nextafter(x, math.inf if n > 0 else -math.inf)
    return x


def ulp_diff(a, b):
    """Number of binary64 ulps from a to b (b - a in representable steps)."""
    ia = struct.unpack("q", struct.pack("d", a))[0]
    ib = struct.unpack("q", struct.pack("d", b))[0]
    return ib - ia


def crt_weights(N):
    M = math.prod(MODULI[:N])
    w = []
    for m in MODULI[:N]:
        Mi = M // m
        w.append(Mi * pow(Mi % m, -1, m))
    return M, w


def build(deltas):
    t = {}
    t["mod"] = MODULI
    t["rcp64"] = [rn_fraction(Fraction(1, m)) for m in MODULI]
    t["rcp32"] = [f32(rn_fraction(Fraction(1, m), 24)) for m in MODULI]
    t["M_hi"], t["M_lo"], t["invM"], t["log2M_fast"], t["log2M_acc"] = [], [], [], [], []
    t["w1"], t["w2hi"], t["w2lo"] = [], [], []
    for N in range(2, 21):
        M, w = crt_weights(N)
        hi = float(rn_fraction(Fraction(M)))
        lo = float(rn_fraction(Fraction(M - int(hi)))) if M != int(hi) else 0.0
        t["M_hi"].append(hi)
        t["M_lo"].append(lo)
        t["invM"].append(rn_fraction(Fraction(1, M)))
        l2lo, l2hi = log2_bracket(M - 1)
        t["log2M_acc"].append(dec9_float32(l2lo / 2 - Fraction(1, 2), l2hi / 2 - Fraction(1, 2)))
        t["log2M_fast"].append(rd_float32(l2lo / 2 - Fraction(3, 2), l2hi / 2 - Fraction(3, 2)))
        row1 = [rn_fraction(Fraction(x)) for x in w]
        d1 = deltas.get("w1", {}).get(str(N), [0] * N)
        t["w1"].append([ulp_add(v, d) for v, d in zip(row1, d1)] + [0.0] * (20 - N))
        if N >= 8:
            g = M.bit_length() - 44 + (N - 1).bit_length()
            his = [(x >> g) << g for x in w]
            los = [rn_fraction(Fraction(x - h)) for x, h in zip(w, his)]
            d2 = deltas.get("w2lo", {}).get(str(N), [0] * N)
            t["w2hi"].append([float(h) for h in his] + [0.0] * (20 - N))
            t["w2lo"].append([ulp_add(v, d) for v, d in zip(los, d2)] + [0.0] * (20 - N))
            for h in his:
                assert float(h) == h  # hi must be exactly representable
    return t


# ---------------------------------------------------------------------------------------------
# reference cross-check (build container only)
# ---------------------------------------------------------------------------------------------
def parse_reference(path):
    src = open(path).read()

    def rows(name, depth_rows):
        i = src.index(name)
        j = src.index("{", i)
        depth, cur, out = 0, None, []
        for k in range(j, len(src)):
            c = src[k]
            if c == "{":
                depth += 1
                if depth == depth_rows:
                    cur = k
            elif c == "}":
                if depth == depth_rows:
                    out.append(src[cur + 1:k])
                depth -= 1
                if depth == 0:
                    break
        return out

    def nums(s):
        pat = r"-?0x[0-9a-fA-F.]+p[-+]?\d+|-?\d+\.\d+e[-+]\d+|-?\d+\.\d+|-?\d+"
        return [float.fromhex(x) if "x" in x else float(x) for x in re.findall(pat, s.replace("F", ""))]

    ref = {}
    mod = [nums(r) for r in rows("constexpr tab_t<double> moduli[20]", 2)]
    ref["mod"] = [int(-r[0]) for r in mod]
    ref["rcp64"] = [r[1] for r in mod]
    ref["rcp32"] = [r[3] for r in mod]
    Mrows = [nums(r) for r in rows("constexpr double M[19][2]", 2)]
    ref["M_hi"] = [r[0] for r in Mrows]
    ref["M_lo"] = [r[1] for r in Mrows]
    ref["invM"] = nums(rows("constexpr double invM[19]", 1)[0])
    i8 = src.index("namespace int8tc")
    vn = src.index("namespace vecnorm")
    ref["log2M_acc"] = [f32(x) for x in nums(rows("constexpr float log2M", 1)[0] if False else src[src.index("{", src.index("constexpr float log2M", i8)):src.index("}", src.index("constexpr float log2M", i8))])]
    ref["log2M_fast"] = [f32(x) for x in nums(src[src.index("{", src.index("constexpr float log2M", vn)):src.index("}", src.index("constexpr float log2M", vn))])]
    ref["w1"] = [nums(r) for r in rows("constexpr double NMi_1[19][20]", 2)]
    w2 = [nums(r) for r in rows("constexpr double NMi_2[13][20][2]", 2)]
    ref["w2hi"] = [r[0::2] for r in w2]
    ref["w2lo"] = [r[1::2] for r in w2]
    return ref


def check_ref(path):
    ref = parse_reference(path)
    base = build({})
    deltas = {"w1": {}, "w2lo": {}}
    problems = []
    for key in ("mod", "rcp64", "rcp32", "M_hi", "M_lo", "invM", "log2M_acc", "log2M_fast"):
        if list(base[key]) != list(ref[key]):
            bad = [(i, a, b) for i, (a, b) in enumerate(zip(base[key], ref[key])) if a != b]
            problems.append((key, bad))
    for N in range(2, 21):
        d = [ulp_diff(a, b) for a, b in zip(base["w1"][N - 2][:N], ref["w1"][N - 2][:N])]
        if any(d):
            deltas["w1"][str(N)] = d
        if N >= 8:
            if base["w2hi"][N - 8][:N] != ref["w2hi"][N - 8][:N]:
                problems.append(("w2hi", N))
            d = [ulp_diff(a, b) for a, b in zip(base["w2lo"][N - 8][:N], ref["w2lo"][N - 8][:N])]
            if any(d):
                deltas["w2lo"][str(N)] = d
    return problems, deltas


# ---------------------------------------------------------------------------------------------
def emit(t, path):
    def arr(name, ctype, vals, fmt):
        body = ", ".join(fmt(v) for v in vals)
        return f"OZ_TABLE({ctype}, {name}, [{len(vals)}]) = {{{body}}};\n"

    def arr2(name, ctype, rows_, fmt):
        body = ",\n".join("  {" + ", ".join(fmt(v) for v in r) + "}" for r in rows_)
        return f"OZ_TABLE({ctype}, {name}, [{len(rows_)}][{len(rows_[0])}]) = {{\n{body}\n}};\n"

    hx = lambda v: float(v).hex() if v != 0 else "0.0"
    hxf = lambda v: (float(v).hex() + "f") if v != 0 else "0.0f"
    s = ("// Generated by tools/gen_tables.py -- do not edit.  Rows are indexed by num_moduli - 2 (or - 8).\n"
         "// The includer defines OZ_TABLE(type, name, dims), e.g. `static const type name dims`.\n")
    s += arr("OZ_MOD", "int", t["mod"], str)
    s += arr("OZ_RCP64", "double", t["rcp64"], hx)
    s += arr("OZ_RCP32", "float", t["rcp32"], hxf)
    # 2^32 mod m_j, symmetric representative: joins the two halves of values beyond 2^57 in the encoder
    s += arr("OZ_POW32", "int", [((1 << 32) % m) - (m if ((1 << 32) % m) > m // 2 else 0) for m in t["mod"]], str)
    # 2^44 mod m_j, symmetric, as a double: values beyond 2^57 are split as h * 2^44 + l and folded as h * c + l (exact, < 2^53)
    s += arr("OZ_POW44", "double", [float(((1 << 44) % m) - (m if ((1 << 44) % m) > m // 2 else 0)) for m in t["mod"]], hx)
    # epilogue Barrett constants (oz_tcgen05.cuh: reduce_mod_u): floor(2^32 / m_j), 2^32 - m_j, and the multiple of m_j that
    # shifts any product |x| <= 2^17 * 127^2 into the unsigned range
    s += arr("OZ_BARRETT_INV", "unsigned", [(1 << 32) // m for m in t["mod"]], lambda v: f"{v}u")
    s += arr("OZ_BARRETT_NEGM", "unsigned", [(1 << 32) - m for m in t["mod"]], lambda v: f"{v}u")
    s += arr("OZ_BARRETT_OFF", "unsigned", [m * (((1 << 17) * 127 * 127 + m - 1) // m) for m in t["mod"]], lambda v: f"{v}u")
    s += arr("OZ_M_HI", "double", t["M_hi"], hx)
    s += arr("OZ_M_LO", "double", t["M_lo"], hx)
    s += arr("OZ_INV_M", "double", t["invM"], hx)
    s += arr("OZ_LOG2M_FAST", "float", t["log2M_fast"], hxf)
    s += arr("OZ_LOG2M_ACC", "float", t["log2M_acc"], hxf)
    s += arr2("OZ_W1", "double", t["w1"], hx)
    s += arr2("OZ_W2_HI", "double", t["w2hi"], hx)
    s += arr2("OZ_W2_LO", "double", t["w2lo"], hx)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    open(path, "w").write(s)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check-ref", default=None, help="path to the reference table.hpp (build container only)")
    ap.add_argument("--out", default=OUT)
    args = ap.parse_args()
    if args.check_ref:
        problems, deltas = check_ref(args.check_ref)
        if problems:
            raise SystemExit(f"derived tables disagree with the reference: {problems}")
        json.dump(deltas, open(DELTAS, "w"), indent=0, sort_keys=True)
        n = sum(sum(1 for x in v if x) for grp in deltas.values() for v in grp.values())
        mx = max((abs(x) for grp in deltas.values() for v in grp.values() for x in v), default=0)
        print(f"reference check OK; {n} entries carry a non-zero ulp offset (max |offset| = {mx})")
    deltas = json.load(open(DELTAS)) if os.path.exists(DELTAS) else {}
    emit(build(deltas), args.out)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
