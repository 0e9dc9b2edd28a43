"""Host restatement of the quotient-estimate reduction of the Shoup kernels (csrc/modarith.cuh canon_est, constants
from b200_ctx_create in csrc/b200ring.cu): for every x < 16q the estimate k never exceeds floor(x/q) and is at most one
short of it, so x - k*q lies in [0, 2q) and one conditional subtraction (or, for FINAL = 2, x - (k+1)*q in [-q, q) and
one sign-based add-back) gives the canonical word.  Random and extreme moduli, extreme x."""
import random


def consts(q):
    bl = q.bit_length()
    sh = bl - 28 if bl > 28 else 0
    qs = q >> sh
    return sh, (1 << 32) // (qs + 1)


def estimate(x, q, sh, mul, plus_one=False):
    xs = x >> sh
    assert xs < 1 << 32, "xs must fit 32 bits for x < 16q"
    return ((xs * mul) >> 32) + (1 if plus_one else 0)


def check(q):
    sh, mul = consts(q)
    assert 16 <= mul <= 32 or q.bit_length() <= 28
    xs = [0, 1, q - 1, q, q + 1, 16 * q - 1, 8 * q - 1, 8 * q, 6 * q - 2]
    for k in range(1, 16):
        xs += [k * q - 1, k * q, k * q + 1, k * q + (q >> 1)]
    xs += [random.randrange(16 * q) for _ in range(200)]
    for x in xs:
        if x >= 16 * q or x >= 1 << 64:
            continue
        k = estimate(x, q, sh, mul)
        t = x // q
        assert t - 1 <= k <= t, (hex(q), hex(x), k, t)
        y = x - k * q
        assert 0 <= y < 2 * q
        y2 = x - estimate(x, q, sh, mul, plus_one=True) * q          # FINAL == 2: one more q up front, add-back if negative
        assert -q <= y2 < q
        assert (y2 + q if y2 < 0 else y2) == x % q


def test_canon_est_bound_random_and_extreme_moduli():
    random.seed(20261019)
    qs = [0x1fffffffffe00001, 0x1000000000248001, 0x10000000042b0001, (1 << 61) - 1, (1 << 60) + 1, (1 << 60) - 1,
          (1 << 31) - 1, (1 << 31) + 11, 0x7fffffff380001, 576460752303439873, (1 << 29) + 1, (1 << 28) - 1, 97]
    for bits in range(29, 62):
        for _ in range(8):
            qs.append(random.randrange(1 << (bits - 1), 1 << bits) | 1)
        qs.append((1 << bits) - 1)
        qs.append((1 << (bits - 1)) + 1)
    for q in qs:
        if q < (1 << 61):
            check(q)
