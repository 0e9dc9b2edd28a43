"""Modulus / root search (ring/primes.go, ring/primes_3n.go) through the C ABI: host-only entry points, so they run
without a GPU.  Checked against the oracle's restatement, the values SURVEY.md section 8(d) lists, plain-Python
restatements of the Go loops, and the properties the reference's own tests assert (ring/ring_test.go:195-245,
ring/ring_3n_test.go:42-73, ring/ntt_3n_test.go:17-36)."""
import math

import pytest

import b200ring as R
from oracle import oracle as O

C3_FIRST, C3_LAST = 0x1000000000248001, 0x10000000042b0001


def _is_prime(n):   # deterministic Miller-Rabin, independent of the library
    if n < 2:
        return False
    for p in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if n % p == 0:
            return n == p
    d, s = n - 1, 0
    while d % 2 == 0:
        d, s = d // 2, s + 1
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


def test_is_prime():
    for x in [0, 1, 2, 3, 4, 561, 1117, 2**61 - 1, 2**61 + 1, C3_FIRST, C3_FIRST + 2, 0xffffffffffffffc5, 0xffffffffffffffff,
              3215031751, 3825123056546413051, 318665857834031151167461 % 2**64]:
        assert R.IsPrime(x) == _is_prime(x), x


@pytest.mark.parametrize("N,bits,count", [(12, 20, 5), (12288, 60, 1), (98304, 60, 24), (3 * 2**13, 45, 8), (36, 30, 6)])
def test_find_3n_rns_primes(N, bits, count):
    got = R.Find3NRNSPrimes(N, bits, count, 100000)
    assert got == [int(q) for q in O.find_3n_rns_primes(N, bits, count)]
    assert got == sorted(set(got)) and all(q % (3 * N) == 1 and _is_prime(q) and q > 2**bits for q in got)
    # nothing skipped: every candidate between consecutive results is composite
    c = ((2**bits - 1) // (3 * N) + 1) * 3 * N + 1
    for q in got[:3]:
        while c < q:
            assert not _is_prime(c)
            c += 3 * N
        c += 3 * N


def test_survey_config_primes():
    qs = R.Find3NRNSPrimes(98304, 60, 24, 1000)
    assert qs[0] == C3_FIRST and qs[-1] == C3_LAST and qs[4] == 0x1000000001568001
    assert R.Find3NRNSPrimes(12288, 60, 1, 1000) == [0x10000000000a1001]
    assert R.Find3NFriendlyPrime(12288, 60) == 0x10000000000a1001


def test_find_3n_rns_primes_errors():
    for args in [(0, 60, 1, 10), (12, 2, 1, 10), (12, 60, 0, 10)]:
        with pytest.raises(ValueError, match="invalid args"):
            R.Find3NRNSPrimes(*args)
    with pytest.raises(ValueError, match=r"could not find enough 3N-friendly primes \(found 1/24\)"):
        R.Find3NRNSPrimes(98304, 60, 24, 10)


def test_find_3n_friendly_prime():
    assert R.Find3NFriendlyPrime(12, 6) == 73   # 3N = 36: 73 = 2*36 + 1 is the first prime = 1 mod 36 above 64
    for N, bits in [(3 * 2**20, 50), (96, 30), (2**4 * 27, 40)]:
        q = R.Find3NFriendlyPrime(N, bits)
        assert q == R.Find3NRNSPrimes(N, bits, 1, 1000)[0] and q % (3 * N) == 1 and _is_prime(q)
    with pytest.raises(ValueError):
        R.Find3NFriendlyPrime(0, 50)


def test_3n_primitive_root_and_decomposition():   # ring/ring_3n_test.go:42-73
    a, b, p = 2, 1, 1117
    m = 3 * (1 << a) * 3
    omega = R.Find3NPrimitiveRoot(p, m)
    assert pow(omega, m, p) == 1 and R.OrderMod(omega, p) == m
    w, w1, w2 = R.Derive3NRoots(p, a, b, omega)
    assert w == omega and w1 == pow(omega, 3 ** (b + 1), p) and w2 == pow(omega, 1 << a, p)
    assert R.OrderMod(w1, p) == 1 << a and R.OrderMod(w2, p) == R.Uint64Pow(3, b + 1)
    with pytest.raises(ValueError, match="omega does not have order 3N"):
        R.Derive3NRoots(p, a, b, w1)
    with pytest.raises(ValueError, match="does not support primitive 3N-th root"):
        R.Derive3NRoots(1013, a, b, omega)
    with pytest.raises(ValueError, match="invalid a,b"):
        R.Derive3NRoots(p, 0, b, omega)


def test_find_primitive_root_of_unity():
    for q, m in [(1117, 36), (0x10000000000a1001, 3 * 12288), (C3_FIRST, 3 * 98304), (0x1fffffffffe00001, 1 << 17)]:
        assert R.HasPrimitiveMthRoot(q, m)
        g = R.PrimitiveRoot(q)
        assert g == int(O.primitive_root(q)) if hasattr(O, "primitive_root") else g >= 3
        assert R.FindPrimitiveRootOfUnity(q, m) == pow(g, (q - 1) // m, q)
        for seed in (1, 2, 0xdeadbeef):
            w = R.FindPrimitiveRootOfUnity(q, m, seed)
            assert R.OrderMod(w, q) == m and pow(w, m, q) == 1 and all(pow(w, m // f, q) != 1 for f in (2, 3) if m % f == 0)
        assert R.FindPrimitiveRootOfUnity(q, m, 7) == R.FindPrimitiveRootOfUnity(q, m, 7)
    assert not R.HasPrimitiveMthRoot(1117, 5)
    with pytest.raises(ValueError, match="no primitive m-th root exists"):
        R.FindPrimitiveRootOfUnity(1117, 5)


def test_order_mod():
    p = 1117
    for g in (2, 3, 5, 962, 1116):
        k = next(k for k in range(1, p) if pow(g, k, p) == 1)
        assert R.OrderMod(g, p) == k
    with pytest.raises(ValueError):
        R.OrderMod(5, 1116)


def _go_generator(bits, nth, direction, k):
    """Plain restatement of ring/primes.go:69-277 for one call on a fresh generator."""
    nxt = prv = (1 << bits) + 1
    cn, cp = not nxt > 2**64 - 1 - nth, not prv < nth
    prv -= nth
    out = []
    while len(out) < k:
        if not (cn or cp) or (direction > 0 and not cn) or (direction < 0 and not cp):
            return out, True
        if cn and direction >= 0:
            if math.log2(float(nxt)) - bits >= 0.5 or nxt > 2**64 - 1 - nth:
                cn = False
            elif _is_prime(nxt):
                out.append(nxt)
                nxt += nth
                if direction == 0:
                    continue
            else:
                nxt += nth
        if cp and direction <= 0 and len(out) < k:
            if bits - math.log2(float(prv)) >= 0.5 or prv < nth:
                cp = False
            elif _is_prime(prv):
                out.append(prv)
                prv -= nth
            else:
                prv -= nth
    return out, False


@pytest.mark.parametrize("bits,nth", [(55, 1 << 17), (61, 1 << 17), (30, 1 << 13), (45, 1 << 15), (20, 1 << 12)])
def test_ntt_friendly_primes_generator(bits, nth):
    for direction, name in ((0, "NextAlternatingPrimes"), (1, "NextUpstreamPrimes"), (-1, "NextDownstreamPrimes")):
        g = R.NewNTTFriendlyPrimesGenerator(bits, nth)
        want, exhausted = _go_generator(bits, nth, direction, 12)
        if exhausted:
            with pytest.raises(ValueError, match="exhausted"):
                getattr(g, name)(12)
            continue
        got = getattr(g, name)(12)
        assert got == want
        assert all(q % nth == 1 and _is_prime(q) for q in got) and len(set(got)) == 12


def test_generate_ntt_primes_like_the_reference():   # ring/ring_test.go:195-245 (testGenerateNTTPrimes)
    nth = 1 << 17
    g = R.NewNTTFriendlyPrimesGenerator(55, nth)
    primes = g.NextAlternatingPrimes(14)
    assert len(set(primes)) == 14 and all(q & (nth - 1) == 1 and R.IsPrime(q) for q in primes)
    assert primes[:3] == [0x80000000080001, 0x80000000440001, 0x7fffffffba0001]   # the 55-bit moduli of Lattigo's examples
    up = g.NextUpstreamPrimes(14)
    assert all(R.IsPrime(q) for q in up) and up == sorted(up) and up[0] > max(primes)
    down = g.NextDownstreamPrimes(14)
    assert all(R.IsPrime(q) for q in down) and down == sorted(down, reverse=True) and down[0] < min(primes)
    assert g.NextAlternatingPrime() not in primes + up + down


def test_generator_exhaustion():
    g = R.NewNTTFriendlyPrimesGenerator(14, 1 << 12)   # 2^14 + k*2^12 + 1 below 2^14.5: a handful of candidates
    with pytest.raises(ValueError, match="cannot NextUpstreamPrime"):
        g.NextUpstreamPrimes(50)
    with pytest.raises(ValueError, match="cannot NextUpstreamPrime"):   # the Go code would spin here
        g.NextUpstreamPrime()
    with pytest.raises(ValueError, match="cannot NextDownstreamPrime"):
        g.NextDownstreamPrimes(50)
    with pytest.raises(ValueError, match="cannot NextAlternatingPrime"):
        g.NextAlternatingPrimes(50)


def test_ring_from_generated_moduli_is_accepted_by_the_constructor_checks():
    """The generated chain passes the constructor's validation (distinct, prime, = 1 mod NthRoot); without a GPU the
    constructor stops at the device, never at the moduli."""
    qs = R.Find3NRNSPrimes(96, 40, 3, 10000)
    try:
        R.NewRing(96, qs)
    except (R.B200Error, ValueError) as e:
        assert getattr(e, "code", 0) == -2, e
