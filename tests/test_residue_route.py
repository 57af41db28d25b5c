"""The exactness argument of the CUDA encoders' short residue routes (mixed-gemmul8_b200/csrc/oz_residue.cuh), checked on
the CPU against integer arithmetic: tests/cxx/residue_model.c restates the device routes operation by operation (same
tables, fma_rd = fma under FE_DOWNWARD) and must return the symmetric residue -- what the reference's mod_8i chain
returns (tests/test_oracle.py proves that half) -- for every magnitude each route is used for.  On the GPU the same
routes are compared byte for byte with the reference library (tests/test_parity_gpu.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cxx", "residue_model.c")
LIB = os.path.join(HERE, "cxx", "libresidue_model.so")
MODS = [256, 255, 253, 251, 247, 241, 239, 233, 229, 227, 223, 217, 211, 199, 197, 193, 191, 181, 179, 173]


@pytest.fixture(scope="module")
def model():
    tables = os.path.join(HERE, "..", "mixed-gemmul8_b200", "csrc", "oz_tables.inc")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(tables)):
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.run([cc, "-O1", "-std=gnu11", "-frounding-math", "-ffp-contract=off", "-fPIC", "-shared", "-o", LIB, SRC, "-lm"], check=True)
    L = C.CDLL(LIB)
    L.residue_model_small_d.argtypes = [C.c_double, C.c_uint]
    L.residue_model_small_f.argtypes = [C.c_float, C.c_uint]
    L.residue_model_big_d.argtypes = [C.c_double, C.c_uint]
    assert [L.residue_model_modulus(j) for j in range(20)] == MODS
    return L


def canon(a, m):
    r = a % m
    if r > m // 2:
        r -= m
    return -128 if (m == 256 and r == 128) else r


def exact_ints(rng, bits_mant, max_bits, per_exp):
    """integers exactly representable with a bits_mant-bit mantissa, magnitudes 1 .. 2^max_bits, both signs"""
    for ebits in range(1, max_bits + 1):
        for _ in range(per_exp):
            nb = min(ebits, bits_mant)
            mant = int(rng.integers(1 << (nb - 1), 1 << nb))
            a = mant << (ebits - nb)
            yield -a if rng.random() < 0.5 else a


def test_small_route_fp64_random(model):
    rng = np.random.default_rng(11)
    for a in exact_ints(rng, 53, 57, 150):                     # the route is used for |a| < 2^57
        if abs(a) >= 1 << 57:
            continue
        for j in range(20):
            assert model.residue_model_small_d(float(a), j) == canon(a, MODS[j]), (a, MODS[j])


def test_small_route_fp64_boundaries(model):
    """quotient boundaries (a = q m + r with r near 0, m/2, m) at the top of the range, where 1/m's rounding error is largest"""
    rng = np.random.default_rng(12)
    for j in range(20):
        m = MODS[j]
        for ebits in (20, 40, 52, 53, 55, 56, 57):
            for _ in range(300):
                q = int(rng.integers(1 << (ebits - 9), (1 << ebits) // m))
                for r in (0, 1, 2, m // 2 - 1, m // 2, m // 2 + 1, m - 2, m - 1):
                    a = q * m + r
                    a = -a if rng.random() < 0.5 else a
                    if abs(a) >= 1 << 57 or int(float(a)) != a:
                        continue
                    assert model.residue_model_small_d(float(a), j) == canon(a, m), (a, m)


def test_small_route_fp32(model):
    rng = np.random.default_rng(13)
    for a in exact_ints(rng, 24, 24, 400):
        if abs(a) >= 1 << 24:
            continue
        for j in range(20):
            assert model.residue_model_small_f(float(a), j) == canon(a, MODS[j]), (a, MODS[j])
    for j in range(20):                                           # every fp32 integer near the top of the range
        m = MODS[j]
        for a in list(range((1 << 24) - 3 * m, 1 << 24)) + list(range(-(1 << 24) + 1, -(1 << 24) + 3 * m)):
            assert model.residue_model_small_f(float(a), j) == canon(a, m), (a, m)


def test_split_route_up_to_2_89(model):
    rng = np.random.default_rng(14)
    for a in exact_ints(rng, 53, 89, 120):                     # 16+ moduli: scaled values reach 2^79; the route holds to 2^89
        for j in range(20):
            assert model.residue_model_big_d(float(a), j) == canon(a, MODS[j]), (a, MODS[j])
