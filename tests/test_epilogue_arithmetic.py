"""The integer arithmetic of the GEMM epilogue (mixed-gemmul8_b200/csrc/oz_gemm.cu), modelled in numpy with the same
32-bit operations and checked against exact integers over the whole range an accumulator can take:
  reduce_mod_u  (CTA-pair kernel)  : shift by a multiple of m into the unsigned range, q = umulhi(x', floor(2^32/m)), one correction
  reduce_mod    (single-CTA kernel): q = mulhi(x, floor(2^32/m)) on the signed value, two corrections
  combine_word / fold_lanes        : four residues per 32-bit word combined in two 16-bit lanes (complex passes)
The accumulator of modulus j > 0 is bounded by k * 127^2 <= 2^17 * 127^2 (symmetric residues of an odd modulus <= 255 lie in
[-127, 127], k <= 2^17); modulus 256 takes the low byte."""
import numpy as np
import pytest

MODS = [256, 255, 253, 251, 247, 241, 239, 233, 229, 227, 223, 217, 211, 199, 197, 193, 191, 181, 179, 173]
KMAX = 2114060288          # 2^17 * 127^2 (kMaxAbsProduct)


def samples(rng, lo, hi, n):
    edge = np.array([lo, lo + 1, -1, 0, 1, hi - 1, hi], dtype=np.int64)
    pw = np.array([s * (1 << b) + d for b in range(31) for s in (-1, 1) for d in (-1, 0, 1)], dtype=np.int64)
    x = np.concatenate([rng.integers(lo, hi + 1, n, dtype=np.int64), edge, pw[(pw >= lo) & (pw <= hi)]])
    return x


@pytest.mark.parametrize("m", MODS[1:])
def test_unsigned_barrett_one_correction(m):
    rng = np.random.default_rng(m)
    inv = (1 << 32) // m
    off = m * ((KMAX + m - 1) // m)
    assert off % m == 0 and off + KMAX < (1 << 32)                       # x + off never wraps
    x = samples(rng, -KMAX, KMAX, 2_000_000)
    near = (rng.integers(-KMAX // m, KMAX // m, 200_000, dtype=np.int64) * m)[:, None] + np.arange(-2, 3)[None, :]   # around multiples of m
    x = np.concatenate([x, near.ravel()])
    x = x[(x >= -KMAX) & (x <= KMAX)]
    xu = (x + off).astype(np.uint64)                                     # uint32 value (no wrap, asserted above)
    q = (xu * np.uint64(inv)) >> np.uint64(32)                           # umulhi
    r = (xu - q * np.uint64(m)).astype(np.int64)
    assert r.min() >= 0 and r.max() < 2 * m                              # one conditional subtraction is enough
    r = np.where(r >= m, r - m, r)
    assert np.array_equal(r, x % m)


@pytest.mark.parametrize("m", MODS[1:])
def test_signed_barrett_two_corrections(m):
    rng = np.random.default_rng(1000 + m)
    inv = (1 << 32) // m
    x = samples(rng, -(1 << 31), (1 << 31) - 1, 2_000_000)
    q = (x * inv) >> 32                                                  # mulhi.s32 (arithmetic shift = floor)
    r = x - q * m
    assert r.min() >= -m and r.max() < 2 * m
    r = np.where(r >= m, r - m, r)
    r = np.where(r < 0, r + m, r)
    assert np.array_equal(r, x % m)


def fold_lanes(t, m, k15):
    ge = ((t + k15) >> 15) & 0x00010001
    return (t - ge * m) & 0xFFFFFFFF


def combine_word(rc, rnew, old, m):
    ml, k15 = (m * 0x00010001) & 0xFFFFFFFF, ((0x8000 - m) * 0x00010001) & 0xFFFFFFFF
    r0, r1 = rnew & 0x00FF00FF, (rnew >> 8) & 0x00FF00FF
    o0, o1 = old & 0x00FF00FF, (old >> 8) & 0x00FF00FF
    aux = np.zeros_like(rnew)
    if rc == 1:
        t0, t1 = o0 + r0, o1 + r1
    elif rc == 2:
        t0, t1 = o0 + ml - r0, o1 + ml - r1
    elif rc == 3:
        t0, t1 = r0 + ml - o0, r1 + ml - o1
    else:
        aux = fold_lanes(o0 + r0, m, k15) | (fold_lanes(o1 + r1, m, k15) << 8)
        t0, t1 = o0 + ml - r0, o1 + ml - r1
    return (fold_lanes(t0 & 0xFFFFFFFF, m, k15) | (fold_lanes(t1 & 0xFFFFFFFF, m, k15) << 8)) & 0xFFFFFFFF, aux & 0xFFFFFFFF


@pytest.mark.parametrize("m", MODS)
def test_packed_residue_combine(m):
    """RC_ADD / RC_SUB / RC_RSUB / RC_KARATSUBA_F on four residues per word against per-byte modular arithmetic"""
    rng = np.random.default_rng(2000 + m)
    a = rng.integers(0, m, (300_000, 4), dtype=np.int64)
    b = rng.integers(0, m, (300_000, 4), dtype=np.int64)
    a[:8] = [[0] * 4, [m - 1] * 4, [0] * 4, [m - 1] * 4, [1] * 4, [m // 2] * 4, [m - 1, 0, 1, m // 2], [0, m - 1, m // 2, 1]]
    b[:8] = [[0] * 4, [m - 1] * 4, [m - 1] * 4, [0] * 4, [m - 1] * 4, [m // 2] * 4, [0, m - 1, m - 1, m // 2], [m - 1, 0, m // 2, m - 1]]
    pack = lambda v: (v[:, 0] | (v[:, 1] << 8) | (v[:, 2] << 16) | (v[:, 3] << 24)).astype(np.int64)
    unpack = lambda w: np.stack([(w >> s) & 0xFF for s in (0, 8, 16, 24)], axis=1)
    new, old = pack(a), pack(b)
    for rc, want in ((1, (b + a) % m), (2, (b - a) % m), (3, (a - b) % m)):
        got, _ = combine_word(rc, new, old, m)
        assert np.array_equal(unpack(got), want), rc
    got, aux = combine_word(4, new, old, m)
    assert np.array_equal(unpack(got), (b - a) % m) and np.array_equal(unpack(aux), (b + a) % m)
