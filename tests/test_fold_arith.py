"""The fold (Solinas) arithmetic of csrc/solinas.cuh restated on Python integers with the device's word semantics
(signed 64-bit accumulators, 32-bit halves), for the primes the fast path admits (b200ring.cu fold_prime_ok: q = 2^k + c,
50 <= k <= 60, |c| <= 2^27).  Checked on random and extreme operands:

  * sol_mul:   result == y*w (mod q), no 64-bit overflow in A, B or the sum, and the window (-0.6, 1.6) * 2^k that
               tools/fold_ranges.py builds every schedule on, for every |y| <= 7.5 * 2^k;
  * sol_fold / sol_fold_u:  same residue, the documented windows;
  * sol_canon: the canonical representative for every signed word with |v| < 2^63 - q;
  * sol_half:  v/2 (mod q) exactly;
  * the table entries {w_lo, w_hi, v_lo, v_hi} of fold_entry fit 32-bit halves.
No GPU: this is the arithmetic the parity tests exercise on the device, stated once where it can be read."""
import random

import pytest

S63 = 1 << 63


def fits64(v):
    return -S63 <= v < S63


def shape(q):
    b = q.bit_length() - 1
    lo = 1 << b
    if b < 63 and q - lo > (lo << 1) - q:
        return b + 1, -((lo << 1) - q)
    return b, q - lo


def split(v, s):
    h = (v + (1 << (s - 1))) >> s
    if not -(1 << 31) <= h < (1 << 31):
        return None                                                  # fold_split refuses: the ring then stays on the Shoup kernels
    return v - (h << s), h


def entry(w, q, k):
    """fold_entry: {w, w*2^32 mod q} as balanced residues split at bit k-32, or None when a half does not fit 32 bits
    (only residues within 2^(k-33) of q/2 when q > 2^k)"""
    v = (w << 32) % q
    wb = w - q if w > q // 2 else w
    vb = v - q if v > q // 2 else v
    a, b = split(wb, k - 32), split(vb, k - 32)
    if a is None or b is None:
        near = lambda x: abs(abs(x) - q // 2) <= (1 << (k - 33)) + 1
        assert near(wb) or near(vb), "only residues next to q/2 may be refused"
        return None
    assert abs(a[0]) <= 1 << (k - 33) and abs(b[0]) <= 1 << (k - 33)
    return a[0], a[1], b[0], b[1]


def sol_mul(y, tw, k, c):
    """sol_mul_sh: y signed 64-bit; returns the signed 64-bit word the device computes, asserting nothing wraps"""
    w_lo, w_hi, v_lo, v_hi = tw
    ylo, yhi = y & 0xffffffff, (y >> 32) & 0xffffffff
    y0 = ylo - (1 << 32) if ylo >> 31 else ylo                       # (int)low word
    y1 = (yhi + (ylo >> 31)) & 0xffffffff
    y1 = y1 - (1 << 32) if y1 >> 31 else y1                          # (int)(high word + carry of the signed low word)
    assert (y1 << 32) + y0 == y
    A = y0 * w_lo + y1 * v_lo
    B = y0 * w_hi + y1 * v_hi
    assert fits64(A) and fits64(B)
    blo, bhi = B & 0xffffffff, B >> 32                               # arithmetic shift of the signed accumulator
    s = k - 32
    sh = blo << s                                                    # pack64(blo << s, blo >> (32 - s)): below 2^k
    r = bhi * (-c) + sh + A
    assert fits64(bhi * (-c)) and fits64(r)
    return r


def sol_fold(v, k, c):
    lo, hi = v & 0xffffffff, v >> 32                                 # hi signed
    s = k - 32
    r = (hi >> s) * (-c) + (((hi & ((1 << s) - 1)) << 32) | lo)
    assert fits64(r)
    return r


def sol_fold_u(v, k, c):
    assert 0 <= v < 1 << 64
    lo, hi = v & 0xffffffff, v >> 32
    s = k - 32
    return (hi >> s) * (-c) + (((hi & ((1 << s) - 1)) << 32) | lo)


def sol_canon(v, q, k, c):
    bias = (S63 // q) * q
    r = sol_fold_u((v + bias) % (1 << 64), k, c)
    if r < 0:
        r += q
    if c < 0 and r >= q:
        r -= q
    return r


def admitted(q):
    """b200ring.cu fold_prime_ok"""
    k, c = shape(q)
    return 50 <= k <= 60 and abs(c) <= 1 << (k - 33)


# 2^60 + c (first and last prime of BASELINE config 3, a power-of-two NTT prime), 2^55 + c, 2^57 - c
PRIMES = [0x1000000000248001, 0x10000000042b0001, 0x1000000000980001, 0x80000000080001, 0x1fffffffffc0001]


def test_admissibility_scales_c_with_k():
    """|c| <= 2^(k-33): the term c*B.hi of a product stays below an eighth of 2^k for every k, as the interval model
    assumes at k = 60.  A 55-bit prime with a 24-bit c is NOT admitted (its products could leave the window)."""
    assert all(admitted(q) for q in PRIMES)
    assert not admitted(0x7fffffff380001) and not admitted(0x7ffffffeac0001)      # 2^55 - c, c > 2^22
    assert not admitted(0x1fffffffffe00001)                                      # Qi60: 2^61 - c, no headroom for signed words
    from bench import MODULI
    assert all(admitted(q) for q in MODULI)


def _is_prime(n):
    if n < 2:
        return False
    for p in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if n % p == 0:
            return n == p
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


@pytest.mark.parametrize("q", [p for p in PRIMES if _is_prime(p)])
def test_fold_arithmetic_on_integers(q):
    rng = random.Random(q)
    k, c = shape(q)
    assert admitted(q)
    unit = 1 << k
    ymax = S63 - (1 << 59)                                            # absolute |y| limit of sol_mul: nothing wraps below it
    ywin = min(ymax, 15 * unit // 2)                                  # what the kernels feed it: |y| <= 7.5 * 2^k (tools/fold_ranges.py Y_MAX)
    ws = [1, 2, q - 1, q // 2, q // 2 + 1, (1 << 32) % q, rng.randrange(q), rng.randrange(q), rng.randrange(q)]
    ys = [0, 1, -1, ymax, -ymax, ymax - 1, ywin, -ywin, (1 << 31), (1 << 31) - 1, -(1 << 31), (1 << 32) - 1, -(1 << 32), unit, -unit, q, -q]
    ys += [rng.randrange(-ymax, ymax + 1) for _ in range(150)] + [rng.randrange(-ywin, ywin + 1) for _ in range(300)]
    for w in ws:
        tw = entry(w, q, k)
        if tw is None:
            continue
        for y in ys:
            if abs(y) > ymax:
                continue
            r = sol_mul(y, tw, k, c)                                     # asserts that no accumulator wraps
            assert (r - y * w) % q == 0
            if abs(y) <= ywin:
                assert -0.6 * unit < r < 1.6 * unit, (hex(q), hex(w), y, r / unit)
            else:
                assert -0.76 * unit < r < 1.76 * unit
    # folds
    for v in [0, 1, -1, S63 - 1, -S63, unit, -unit, q, 7 * unit + 12345, -7 * unit - 999] + [rng.randrange(-S63, S63) for _ in range(300)]:
        f = sol_fold(v, k, c)
        assert (f - v) % q == 0
        h = abs(c) * (1 << (63 - k))
        assert -h <= f <= unit + h
        if abs(v) < S63 - q:
            assert sol_canon(v, q, k, c) == v % q
            half = (v + (q if v & 1 else 0)) >> 1                      # sol_half
            assert (2 * half - v) % q == 0 and fits64(v + q)
    for v in [0, 1, (1 << 64) - 1, S63, S63 - 1, q, 8 * q if 8 * q < (1 << 64) else q] + [rng.randrange(1 << 64) for _ in range(300)]:
        f = sol_fold_u(v, k, c)
        assert (f - v) % q == 0
        assert -(abs(c) << (64 - k)) <= f <= unit + (abs(c) << (64 - k))


def test_every_c3_twiddle_half_fits():
    """the 24 primes of BASELINE config 3: random residues always split into 32-bit halves (fold_entry never refuses them)"""
    from bench import MODULI
    rng = random.Random(3)
    for q in MODULI:
        k, c = shape(q)
        assert k == 60 and 0 < c < 1 << 27
        for _ in range(200):
            assert entry(rng.randrange(q), q, k) is not None
        for w in (1, q - 1, (q - 1) // 3, q // 4):
            assert entry(w, q, k) is not None
        assert entry(q // 2, q, k) is None       # the residue next to q/2 is the one the split cannot hold (c > 0)
