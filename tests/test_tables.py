"""The constant tables are mathematics: check them from first principles (exact integers)."""
import math
import os
import re
import sys
from fractions import Fraction

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_tables


def load_tables():
    txt = open(os.path.join(ROOT, "mixed-gemmul8_b200", "csrc", "oz_tables.inc")).read()
    tabs = {}
    for mobj in re.finditer(r"OZ_TABLE\((\w+), (\w+), ((?:\[\d+\])+)\) = \{(.*?)\};", txt, re.S):
        _, name, dims, body = mobj.groups()
        vals = [float.fromhex(x.rstrip("f")) if "x" in x else float(x.rstrip("f")) for x in re.findall(r"-?0x[0-9a-fA-F.]+p[-+]?\d+f?|-?\d+\.?\d*f?", body)]
        shape = [int(d) for d in re.findall(r"\d+", dims)]
        if len(shape) == 2:
            vals = [vals[i * shape[1]:(i + 1) * shape[1]] for i in range(shape[0])]
        tabs[name] = vals
    return tabs


T = load_tables()
MOD = [int(x) for x in T["OZ_MOD"]]


def test_moduli_pairwise_coprime_and_descending():
    assert MOD[0] == 256 and all(a > b for a, b in zip(MOD, MOD[1:]))
    for i in range(20):
        for j in range(i + 1, 20):
            assert math.gcd(MOD[i], MOD[j]) == 1


def test_reciprocals_are_correctly_rounded():
    for m, r64, r32 in zip(MOD, T["OZ_RCP64"], T["OZ_RCP32"]):
        assert r64 == gen_tables.rn_fraction(Fraction(1, m))
        assert r32 == gen_tables.f32(gen_tables.rn_fraction(Fraction(1, m), 24))


@pytest.mark.parametrize("N", range(2, 21))
def test_M_and_weights(N):
    M = math.prod(MOD[:N])
    hi, lo = T["OZ_M_HI"][N - 2], T["OZ_M_LO"][N - 2]
    assert abs(int(hi) + int(lo) - M) <= abs(M) * 2 ** -100 or Fraction(hi) + Fraction(lo) == M or abs(Fraction(hi) + Fraction(lo) - M) < Fraction(M, 2 ** 104)
    assert abs(T["OZ_INV_M"][N - 2] * M - 1) < 2 ** -50
    # exact CRT weights: w_j = 1 mod m_j, 0 mod the others
    _, w = gen_tables.crt_weights(N)
    for j, wj in enumerate(w):
        for i, m in enumerate(MOD[:N]):
            assert wj % m == (1 if i == j else 0)
        # single-double weight within a few hundred ulps of the exact value (the reference's literals carry that noise)
        assert abs(Fraction(T["OZ_W1"][N - 2][j]) - wj) <= Fraction(wj, 2 ** 45)
        if N >= 8:
            h, l = T["OZ_W2_HI"][N - 8][j], T["OZ_W2_LO"][N - 8][j]
            assert abs(Fraction(h) + Fraction(l) - wj) <= Fraction(wj, 2 ** 85)
            # hi parts share a grid coarse enough that sum_j hi_j * 255 is exact in binary64
            g = M.bit_length() - 44 + (N - 1).bit_length()
            assert int(h) % (1 << g) == 0 and N * 255 * M < (1 << (53 + g))


def test_budgets_monotone():
    for name in ("OZ_LOG2M_FAST", "OZ_LOG2M_ACC"):
        v = T[name]
        assert all(a < b for a, b in zip(v, v[1:]))
    for N in range(2, 21):
        true = math.log2(math.prod(MOD[:N]) - 1) / 2
        assert abs(T["OZ_LOG2M_FAST"][N - 2] - (true - 1.5)) < 1e-5
        assert abs(T["OZ_LOG2M_ACC"][N - 2] - (true - 0.5)) < 1e-5


def test_generator_is_reproducible(tmp_path):
    out = tmp_path / "t.inc"
    import json
    deltas = json.load(open(gen_tables.DELTAS))
    gen_tables.emit(gen_tables.build(deltas), str(out))
    assert out.read_text() == open(os.path.join(ROOT, "mixed-gemmul8_b200", "csrc", "oz_tables.inc")).read()


@pytest.mark.skipif(not os.path.exists("/root/reference/GEMMul8/src/table.hpp"), reason="reference only in the build container")
def test_tables_equal_reference_bit_for_bit():
    problems, deltas = gen_tables.check_ref("/root/reference/GEMMul8/src/table.hpp")
    assert not problems
    import json
    assert deltas == json.load(open(gen_tables.DELTAS))
