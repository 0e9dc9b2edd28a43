"""The remaining operators of ring/operations.go on the GPU: Shift (golden vector of ring/ring_test.go:904-916),
MultByMonomial (property of ring_test.go:886-902 + the Go loops restated word for word), and the scalar operators, whose
per-limb Montgomery scalars are prepared on the host exactly as the Go methods do (ring/operations.go:151-275)."""
import numpy as np
import pytest

from conftest import QI60
import b200ring
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def go_mult_by_monomial(p1, k, q, N):
    """ring/operations.go:306-363, loop for loop."""
    shift = (k + (N << 1)) % (N << 1)
    if shift == 0:
        return p1.copy()
    tmpx = p1.copy() if shift < N else (np.uint64(q) - p1)
    shift %= N
    out = np.empty_like(p1)
    out[:shift] = np.uint64(q) - tmpx[N - shift:]
    out[shift:] = tmpx[: N - shift]
    return out


def rand_poly(seed, batch, moduli, N):
    out = np.empty((batch, len(moduli), N), dtype=np.uint64)
    for b in range(batch):
        for l, q in enumerate(moduli):
            out[b, l] = O.fill_uniform(seed, b, l, q, N)
    return out


def test_shift_golden_vector():
    """ring_test.go:904-916."""
    r = b200ring.NewRing(16, [97])
    p1, p2 = r.NewPoly().set(np.arange(16, dtype=np.uint64).reshape(1, 1, 16)), r.NewPoly()
    r.Shift(p1, 3, p2)
    assert p2.get()[0, 0].tolist() == [3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 0, 1, 2]
    r.Shift(p2, -3, p2)                                                       # in place, negative k
    assert p2.get()[0, 0].tolist() == list(range(16))


@pytest.mark.parametrize("N,kind", [(16, "std"), (1 << 12, "std"), (3 << 10, "m3")])
def test_mult_by_monomial(N, kind):
    Q = QI60[:3] if kind == "std" else O.find_3n_rns_primes(N, 60, 3)
    r = b200ring.NewRing(N, Q)
    host = rand_poly(0xD0, 2, Q, N)
    host[0, 0, 0] = 0                                                         # q - 0 = q is kept, like the Go loops
    p1, t, w = r.NewPoly(2).set(host), r.NewPoly(2), r.NewPoly(2)
    r.MultByMonomial(p1, 1, t); r.MultByMonomial(t, 8, t)                     # ring_test.go:895-900
    r.MultByMonomial(p1, 9, w)
    for l, q in enumerate(Q):
        assert (t.get()[:, l] % np.uint64(q) == w.get()[:, l] % np.uint64(q)).all()
    for k in (0, 1, 9, N - 1, N, N + 5, 2 * N - 1, -3, -N - 2):
        r.MultByMonomial(p1, k, w)
        got = w.get()
        for b in range(2):
            for l, q in enumerate(Q):
                assert (got[b, l] == go_mult_by_monomial(host[b, l], k, q, N)).all(), (k, b, l)
    r.AtLevel(0).MultByMonomial(p1, 5, p1)                                    # in place, one level
    got = p1.get()
    assert (got[:, 0] == np.stack([go_mult_by_monomial(host[b, 0], 5, Q[0], N) for b in range(2)])).all()
    assert (got[:, 1:] == host[:, 1:]).all()
    if kind == "std":                                                         # X^k * p in the NTT domain is a slot-wise product
        sub = O.Pow2SubRing(N, Q[1])
        mono = np.zeros(N, dtype=np.uint64); mono[3] = 1
        want = sub.INTT(O.vec_op("MulCoeffsBarrett", Q[1], sub.NTT(host[1, 1]), sub.NTT(mono)))
        assert (go_mult_by_monomial(host[1, 1], 3, Q[1], N) % np.uint64(Q[1]) == want).all()


def test_scalar_operators_match_big_integers():
    N, Q = 64, QI60[:3]
    r = b200ring.NewRing(N, Q)
    host, acc = rand_poly(0xD1, 1, Q, N), rand_poly(0xD2, 1, Q, N)
    p, o = r.NewPoly().set(host), r.NewPoly()
    sc, big = 0x123456789ABCDEF1, 3 ** 150 + 7
    chk = lambda got, f: all(got[0, l].tolist() == [f(int(x), int(a), q) for x, a in zip(host[0, l], acc[0, l])] for l, q in enumerate(Q))
    r.MulScalar(p, sc, o)
    assert chk(o.get(), lambda x, a, q: x * sc % q)
    o.set(acc); r.MulScalarThenAdd(p, sc, o)
    assert chk(o.get(), lambda x, a, q: (a + x * sc) % q)
    o.set(acc); r.MulScalarThenSub(p, sc, o)
    assert chk(o.get(), lambda x, a, q: (a - x * sc) % q)
    r.MulScalarBigint(p, big, o)
    assert chk(o.get(), lambda x, a, q: x * big % q)
    o.set(acc); r.MulScalarBigintThenAdd(p, big, o)
    assert chk(o.get(), lambda x, a, q: (a + x * big) % q)
    r.AddScalarBigint(p, big, o)
    assert chk(o.get(), lambda x, a, q: (x + big) % q)
    r.SubScalarBigint(p, big, o)
    assert chk(o.get(), lambda x, a, q: (x - big) % q)
    r.MulRNSScalarMontgomery(p, [7 * (1 << 64) % q for q in Q], o)
    assert chk(o.get(), lambda x, a, q: x * 7 % q)
    coeffs = [r.NewPoly().set(rand_poly(0xD3 + i, 1, Q, N)) for i in range(3)]    # p2 = c0 + c1*s + c2*s^2 slot-wise
    r.EvalPolyScalar(coeffs, sc, o)
    c = [x.get() for x in coeffs]
    for l, q in enumerate(Q):
        assert o.get()[0, l].tolist() == [(int(c[0][0, l, j]) + int(c[1][0, l, j]) * sc + int(c[2][0, l, j]) * sc * sc) % q for j in range(N)]


@pytest.mark.parametrize("kind", ["std", "m3"])
def test_div_by_last_modulus_many_like_the_reference(kind):
    """ring/ring_test.go:242-331 as written there: random coefficients below Q/10 set with SetCoefficientsBigint, `level`
    rescales with DivFloor/RoundByLastModulusMany[NTT], compared with big-integer division through PolyToBigint."""
    N = 64 if kind == "std" else 48
    Q = QI60[:4] if kind == "std" else b200ring.Find3NRNSPrimes(N, 60, 4, 100000)
    r = b200ring.NewRing(N, Q)
    level = r.Level()
    rng = np.random.default_rng(17)
    coeffs = [int.from_bytes(rng.bytes(32), "little") % (r.Modulus() // 10) for _ in range(N)]
    p0, p1 = r.NewPoly(), r.NewPoly()
    r.SetCoefficientsBigint(coeffs, p0)
    assert r.PolyToBigint(p0) == coeffs and r.PolyToString(p0)[3] == str(coeffs[3])
    assert r.PolyToBigint(p0, gap=4) == coeffs[::4]
    cent = r.PolyToBigintCentered(p0)
    assert all((c - x) % r.Modulus() == 0 and -(r.Modulus() // 2) - 1 <= c < r.Modulus() // 2 for c, x in zip(cent, coeffs))
    for rnd in (False, True):
        want = list(coeffs)
        for j in range(level):
            ql = Q[level - j]
            want = [(c + (ql - 1) // 2) // ql if rnd else c // ql for c in want]
        (r.DivRoundByLastModulusMany if rnd else r.DivFloorByLastModulusMany)(level, p0, p1)
        assert r.AtLevel(0).PolyToBigint(p1) == [w % Q[0] for w in want]
        pn = r.NewPoly()
        r.NTT(p0, pn)
        (r.DivRoundByLastModulusManyNTT if rnd else r.DivFloorByLastModulusManyNTT)(level, pn, p1)
        r.AtLevel(0).INTT(p1, p1)
        assert r.AtLevel(0).PolyToBigint(p1) == [w % Q[0] for w in want]
    assert r.PolyToBigint(p0) == coeffs                      # the operand is left as it was


def test_ring_accessors_and_monomial():
    N, Q = 32, QI60[:3]
    r = b200ring.NewRing(N, Q)
    assert r.LogN() == 5 and r.Modulus() == Q[0] * Q[1] * Q[2] and r.AtLevel(1).Modulus() == Q[0] * Q[1]
    assert abs(r.LogModuli() - sum(np.log2(float(q)) for q in Q)) < 1e-9
    assert r.MRedConstants() == [pow(q, -1, 1 << 64) for q in Q] and r.BRedConstants() == [((1 << 128) // q >> 64, (1 << 128) // q & (2**64 - 1)) for q in Q]
    assert r.NewRNSScalar() == [0, 0, 0]
    for i in (0, 5, N, N + 3, 2 * N + 1, -1):
        m = r.NewMonomialXi(i).get()[0]
        k = i & (2 * N - 1)
        for l, q in enumerate(Q):
            want = np.zeros(N, dtype=np.uint64)
            want[k % N] = 1 if k < N else q - 1
            assert (m[l] == want).all(), (i, l)
    # X^i through the ring product: MultByMonomial(p, i) == p * NewMonomialXi(i)
    host = rand_poly(0xE1, 1, Q, N)
    p, o, o2 = r.NewPoly().set(host), r.NewPoly(), r.NewPoly()
    r.MultByMonomial(p, 7, o)
    r.MulPoly(p, r.NewMonomialXi(7), o2)
    assert (o.get() == o2.get()).all()
