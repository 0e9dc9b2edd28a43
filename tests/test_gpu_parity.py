"""GPU parity tests: every call goes through the C ABI (b200ring -> libb200ring.so) and is compared
bit-for-bit with the CPU oracle on the same seeded inputs, with the committed golden vectors, and --
at BASELINE's full sizes -- through size-independent properties (round trip, product, linearity,
spot evaluation by Horner).  Mirrors ring/ntt_test.go, ring/ntt_3n_test.go, ring/ring_test.go."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, QI60
import b200ring
from oracle import oracle as O

pytestmark = pytest.mark.gpu

C3_PRIMES = [0x1000000000248001, 0x1000000000b00001, 0x1000000000cf8001, 0x1000000000ef0001,
             0x1000000001568001, 0x10000000019a0001, 0x1000000001be0001, 0x1000000001cb8001,
             0x1000000001eb0001, 0x10000000026d8001, 0x1000000002720001, 0x1000000002888001,
             0x1000000002c30001, 0x1000000002c78001, 0x1000000002eb8001, 0x10000000030b0001,
             0x1000000003260001, 0x1000000003530001, 0x1000000003650001, 0x1000000003698001,
             0x1000000003e30001, 0x1000000003e78001, 0x1000000003ec0001, 0x10000000042b0001]


def rand_poly(seed, batch, moduli, N):
    out = np.empty((batch, len(moduli), N), dtype=np.uint64)
    for b in range(batch):
        for l, q in enumerate(moduli):
            out[b, l] = O.fill_uniform(seed, b, l, q, N)
    return out


# ---------------------------------------------------------------- power-of-two ring
def test_kat_ntt_test_go():
    """ring/ntt_test.go:91-119 on the GPU: NTT(poly) == polyNTT and INTT(NTT(x)) == x, both limbs."""
    kat = json.load(open(os.path.join(GOLDEN, "ntt_kat.json")))["vectors"]
    for v in kat:
        r = b200ring.NewRing(v["N"], v["Qis"])
        assert r.Type() == b200ring.Standard
        x, y, z = r.NewPoly(), r.NewPoly(), r.NewPoly()
        x.set(np.array(v["poly"], dtype=np.uint64))
        y.set(np.array(v["polyNTT"], dtype=np.uint64))
        r.NTT(x, z)
        assert r.Equal(z, y), "transformed poly and polyNTT should match"
        assert (z.get() == y.get()).all()
        r.INTT(z, z)
        assert r.Equal(z, x), "invNTT should reverse NTT"
        assert (z.get() == x.get()).all()


def test_standard_tables_bit_identical_to_subring_go():
    for N in (16, 1024):
        r = b200ring.NewRing(N, [576460752303439873, 576460752303702017][: 2 if N == 16 else 1] if N == 16 else QI60[:2])
        for i, q in enumerate(r.ModuliChain()):
            s = O.Pow2SubRing(N, q)
            sr = r.SubRings[i]
            assert (sr.Modulus, sr.MRedConstant, sr.BRedConstant, sr.PrimitiveRoot, sr.NInv) == \
                   (q, s.MRedConstant, s.BRedConstant, s.PrimitiveRoot, s.NInv)
            rf, rb = s.tables()
            assert (sr.RootsForward == rf).all() and (sr.RootsBackward == rb).all()
            assert sr.Mask == (1 << (q - 1).bit_length()) - 1


@pytest.mark.parametrize("logN", [3, 4, 7, 9, 10, 11, 12, 13, 14, 15, 16, 17])
def test_standard_ntt_matches_oracle(logN):
    N = 1 << logN
    moduli = QI60[:3]
    batch = 2
    r = b200ring.NewRing(N, moduli)
    host = rand_poly(0xB2000002, batch, moduli, N)
    p, o = r.NewPoly(batch).set(host), r.NewPoly(batch)
    r.NTT(p, o)
    got = o.get()
    subs = [O.Pow2SubRing(N, q) for q in moduli]
    for b in range(batch):
        for l in range(len(moduli)):
            assert (got[b, l] == subs[l].NTT(host[b, l])).all(), (logN, b, l)
    r.NTTLazy(p, o)
    lazy = o.get()
    for l, q in enumerate(moduli):
        assert (lazy[:, l] % np.uint64(q) == got[:, l]).all() and int(lazy[:, l].max()) <= 6 * q - 2
    r.INTT(o, o)                       # in place, from lazy input < 4q?  no: INTT wants < 2q -> use canonical
    o.set(got)
    r.INTT(o, o)
    assert (o.get() == host).all()
    o.set(got)
    r.INTTLazy(o, o)
    lz = o.get()
    for l, q in enumerate(moduli):
        assert (lz[:, l] % np.uint64(q) == host[:, l]).all() and int(lz[:, l].max()) <= 2 * q - 1
    # AtLevel: only limbs 0..level are touched (ring/ntt.go:128)
    o.set(host)
    r.AtLevel(0).NTT(o, o)
    after = o.get()
    assert (after[:, 0] == got[:, 0]).all() and (after[:, 1:] == host[:, 1:]).all()


def test_standard_polymul_pipeline_vs_schoolbook():
    """NTT -> MForm -> MulCoeffsMontgomery -> INTT (ring/ring_test.go:106-110) == negacyclic schoolbook."""
    N, moduli = 256, QI60[:2]
    r = b200ring.NewRing(N, moduli)
    ha, hb = rand_poly(21, 1, moduli, N), rand_poly(22, 1, moduli, N)
    a, b, c = r.NewPoly().set(ha), r.NewPoly().set(hb), r.NewPoly()
    r.NTT(a, a); r.NTT(b, b)
    r.MForm(a, a)
    r.MulCoeffsMontgomery(a, b, c)
    r.INTT(c, c)
    want = np.stack([O.schoolbook_negacyclic(N, q, ha[0, l], hb[0, l]) for l, q in enumerate(moduli)])
    assert (c.get()[0] == want).all()
    a.set(ha); b.set(hb)
    r.MulPoly(a, b, c)
    assert (c.get()[0] == want).all()


# ---------------------------------------------------------------- element-wise family
ALL_BINARY = ["Add", "AddLazy", "Sub", "SubLazy", "MulCoeffsBarrett", "MulCoeffsBarrettLazy",
              "MulCoeffsBarrettThenAdd", "MulCoeffsBarrettThenAddLazy", "MulCoeffsMontgomery",
              "MulCoeffsMontgomeryLazy", "MulCoeffsMontgomeryThenAdd", "MulCoeffsMontgomeryThenAddLazy",
              "MulCoeffsMontgomeryLazyThenAddLazy", "MulCoeffsMontgomeryThenSub", "MulCoeffsMontgomeryThenSubLazy",
              "MulCoeffsMontgomeryLazyThenSubLazy", "MulCoeffsMontgomeryLazyThenNeg"]
ALL_UNARY = ["Neg", "Reduce", "ReduceLazy", "MForm", "MFormLazy", "IMForm"]
ALL_SCALAR = ["MulScalarMontgomery", "MulScalarMontgomeryLazy", "MulScalarMontgomeryThenAdd", "AddScalar", "SubScalar"]


@pytest.mark.parametrize("logN", [8, 10, 11, 13])
def test_standard_fused_product_equals_operator_sequence(logN):
    """ring/ring_test.go:106-110 at the sizes where the column-pass kernel family changes shape."""
    N, moduli = 1 << logN, QI60[:3]
    r = b200ring.NewRing(N, moduli)
    ha, hb = rand_poly(51, 2, moduli, N), rand_poly(52, 2, moduli, N)
    a, b, c = r.NewPoly(2).set(ha), r.NewPoly(2).set(hb), r.NewPoly(2)
    r.MulPoly(a, b, c)
    got = c.get()
    r.NTT(a, a); r.NTT(b, b); r.MForm(a, a); r.MulCoeffsMontgomery(a, b, c); r.INTT(c, c)
    assert (c.get() == got).all()
    for l, q in enumerate(moduli[:1]):   # one limb against the negacyclic schoolbook product
        if N <= 1 << 10:
            assert (got[0, l] == O.schoolbook_negacyclic(N, q, ha[0, l], hb[0, l])).all()


@pytest.mark.parametrize("N", [16, 24, 3 * 2**10])
def test_vec_ops_bit_exact(N):
    is3n = N % 3 == 0
    moduli = O.find_3n_rns_primes(N, 60, 3) if is3n else QI60[18:21]
    r = b200ring.NewRing(N, moduli)
    batch = 2
    h1, h2, h3 = (rand_poly(s, batch, moduli, N) for s in (31, 32, 33))
    # edge values of ring/ring_test.go:534-670 in the first slots
    for l, q in enumerate(moduli):
        h1[0, l, :3] = [1, q - 1, 0]
        h2[0, l, :3] = [q - 1, q - 1, 1]
    p1, p2 = r.NewPoly(batch).set(h1), r.NewPoly(batch).set(h2)
    for name in ALL_BINARY:
        p3 = r.NewPoly(batch).set(h3)
        getattr(r, name)(p1, p2, p3)
        got = p3.get()
        for b in range(batch):
            for l, q in enumerate(moduli):
                assert (got[b, l] == O.vec_op(name, q, h1[b, l], h2[b, l], h3[b, l])).all(), (name, b, l)
    big = h1.copy(); big[..., ::2] = np.uint64(2**64 - 1) - big[..., ::2]    # arbitrary 64-bit inputs for Reduce
    for name in ALL_UNARY:
        src = big if name.startswith("Reduce") else h1
        p, o = r.NewPoly(batch).set(src), r.NewPoly(batch)
        getattr(r, name)(p, o)
        got = o.get()
        for b in range(batch):
            for l, q in enumerate(moduli):
                assert (got[b, l] == O.vec_op(name, q, src[b, l])).all(), (name, b, l)
    scal = [12345678901234567 % q for q in moduli]
    for name in ALL_SCALAR:
        p3 = r.NewPoly(batch).set(h3)
        getattr(r, name)(p1, scal, p3)
        got = p3.get()
        for b in range(batch):
            for l, q in enumerate(moduli):
                assert (got[b, l] == O.vec_op(name, q, h1[b, l], None, h3[b, l], scalar=scal[l])).all(), (name, b, l)
    # in place: p1 = p1 * p2
    r.MulCoeffsMontgomery(p1, p2, p1)
    got = p1.get()
    for l, q in enumerate(moduli):
        assert (got[1, l] == O.vec_op("MulCoeffsMontgomery", q, h1[1, l], h2[1, l])).all()


# ---------------------------------------------------------------- 3N (Matrix) ring
def test_3n_prototype_vectors_on_gpu():
    vec = json.load(open(os.path.join(GOLDEN, "ntt3n_proto.json")))["vectors"]
    for v in vec:
        N = v["N"]
        if N < 8:
            continue                       # ring constructors need N >= 8 (ring/ring.go:20-23)
        r = b200ring.NewRing(N, [v["p"]], roots=[v["w"]])
        assert r.Type() == b200ring.Matrix
        x, y = r.NewPoly().set(np.array([[v["coeffs"]]], dtype=np.uint64)), r.NewPoly()
        r.NTT(x, y, tree_order=True)
        assert [int(t) for t in y.get()[0, 0]] == v["ntt_tree_order"]
        r.INTT(y, y, tree_order=True)
        assert (y.get() == x.get()).all()
        r.NTT(x, y)
        assert [int(t) for t in y.get()[0, 0]] == O.py_3n_ntt(N, v["p"], v["w"], v["coeffs"])


@pytest.mark.parametrize("N", [6, 12, 18, 24, 36, 48, 96, 192, 3 * 2**7, 9 * 2**6, 3 * 2**10, 3 * 2**11, 3 * 2**12, 3 * 2**13, 27 * 2**10])
def test_3n_matches_oracle_roundtrip_product(N):
    moduli = O.find_3n_rns_primes(N, 60, 2)
    batch = 2
    r = b200ring.NewRing(N, moduli)               # deterministic omega (subring.go:255-290)
    subs = [O.ThreeNSubRing(N, q) for q in moduli]
    for s, sr in zip(subs, r.SubRings):
        assert s.omega == sr.Root
    host = rand_poly(0xB2000001, batch, moduli, N)
    p, o = r.NewPoly(batch).set(host), r.NewPoly(batch)
    r.NTT(p, o)
    got = o.get()
    for b in range(batch):
        for l in range(len(moduli)):
            assert (got[b, l] == subs[l].NTT(host[b, l])).all(), (N, b, l)
    if N <= 192:
        assert (got[0, 0] == subs[0].NTT_horner(host[0, 0])).all()       # ntt_3n.go:99-108
    r.INTT(o, o)
    assert (o.get() == host).all()                                       # ntt_3n_test.go:139-222
    z = r.NewPoly(batch).set(host)
    for _ in range(8):
        r.NTT(z, z); r.INTT(z, z)
    assert (z.get() == host).all()
    zero, one = np.zeros((batch, len(moduli), N), dtype=np.uint64), np.zeros((batch, len(moduli), N), dtype=np.uint64)
    one[..., 0] = 1
    r.NTT(z.set(zero), z); assert (z.get() == 0).all()
    r.NTT(z.set(one), z); assert (z.get() == 1).all()
    # product == schoolbook mod X^N - X^(N/2) + 1 (ntt_3n_test.go:455-554)
    ha, hb = rand_poly(41, batch, moduli, N), rand_poly(42, batch, moduli, N)
    a, b, c = r.NewPoly(batch).set(ha), r.NewPoly(batch).set(hb), r.NewPoly(batch)
    r.MulPoly(a, b, c)
    got = c.get()
    if N <= 3 * 2**7:
        for bi in range(batch):
            for l, q in enumerate(moduli):
                assert (got[bi, l] == O.schoolbook_3n(N, q, ha[bi, l], hb[bi, l])).all()
    r.NTT(a, a); r.NTT(b, b); r.MForm(a, a); r.MulCoeffsMontgomery(a, b, c); r.INTT(c, c)
    assert (c.get() == got).all()                                        # fused product == operator sequence


@pytest.mark.parametrize("N", [12, 24, 96, 3 * 2**10, 3 * 2**12])
def test_3n_small_prime_cases_of_the_reference_on_gpu(N):
    """ring/ntt_3n_test.go:139-222 with the reference's own parameters on the GPU path: 31-bit primes from
    Find3NRNSPrimes(N, 31, .., 1<<16) (32-bit words, far from the 60/61-bit chains of the other cases: the Shoup kernels with
    their quotient-estimate reduction, resident and two-pass shapes): NTT == oracle, eight Forward+Backward cycles, the zero
    and the all-ones vector, and at the reference's own sizes the definition out[k] = f(omega^E[k])."""
    moduli = O.find_3n_rns_primes(N, 31, 2)
    assert all(q.bit_length() == 32 and q % (3 * N) == 1 for q in moduli)
    r = b200ring.NewRing(N, moduli)
    subs = [O.ThreeNSubRing(N, q) for q in moduli]
    host = rand_poly(31, 2, moduli, N)
    p, o = r.NewPoly(2).set(host), r.NewPoly(2)
    r.NTT(p, o)
    got = o.get()
    for b in range(2):
        for l in range(2):
            assert (got[b, l] == subs[l].NTT(host[b, l])).all(), (N, b, l)
    if N <= 24:
        q, s = moduli[0], subs[0]
        E = [int(e) for e in O.totatives(3 * N)]
        for k in range(N):
            x = pow(s.omega, E[k], q)
            assert int(got[0, 0, k]) == sum(int(c) * pow(x, j, q) for j, c in enumerate(host[0, 0])) % q
    z = r.NewPoly(2).set(host)
    for _ in range(8):
        r.NTT(z, z); r.INTT(z, z)
    assert (z.get() == host).all()
    for v in (np.zeros_like(host), np.ones_like(host)):
        r.NTT(z.set(v), z); r.INTT(z, z)
        assert (z.get() == v).all()
    r.NTT(z.set(np.zeros_like(host)), z)
    assert not z.get().any()


def test_3n_random_omega_from_go_side():
    """the Go transformer draws omega at random; the bridge passes it and slots must follow it."""
    N = 48
    q = O.find_3n_rns_primes(N, 60, 1)[0]
    g = O.lib().ref_primitive_root(q)
    omega = pow(g, 5 * (q - 1) // (3 * N), q)      # another primitive 3N-th root (5 is coprime to 144)
    r = b200ring.NewRing(N, [q], roots=[omega])
    x = rand_poly(5, 1, [q], N)
    y = r.NewPoly()
    r.NTT(r.NewPoly().set(x), y)
    assert [int(t) for t in y.get()[0, 0]] == O.py_3n_ntt(N, q, omega, x[0, 0])
    with pytest.raises(ValueError):
        b200ring.NewRing(N, [q], roots=[pow(omega, 2, q)])   # order 3N/2: not primitive


def test_plugin_granularity_host_slices():
    """ring.NumberTheoreticTransformer (ring/ntt.go:15-22): per-limb host slices, in place allowed."""
    N = 64
    r = b200ring.NewRing(N, [576460752303439873, 576460752303702017])
    x = O.fill_uniform(9, 0, 1, 576460752303702017, N)
    s = O.Pow2SubRing(N, 576460752303702017)
    y = r.SubRings[1].NTT(x)
    assert (y == s.NTT(x)).all()
    buf = x.copy()
    r.SubRings[1].NTT(buf, buf)
    assert (buf == y).all()
    assert (r.SubRings[1].INTT(y) == x).all()
    with pytest.raises(b200ring.B200Error):
        r.SubRings[1].NTT(x[: N - 1])            # short slice: the reference panics (ntt.go:211-214)
    N3 = 24
    q3 = O.find_3n_rns_primes(N3, 60, 1)[0]
    r3 = b200ring.NewRing(N3, [q3])
    x3 = O.fill_uniform(9, 0, 0, q3, N3)
    assert (r3.SubRings[0].NTT(x3) == O.ThreeNSubRing(N3, q3).NTT(x3)).all()


def test_shape_errors_panic_like_reference():
    r = b200ring.NewRing(16, [576460752303439873, 576460752303702017])
    r2 = b200ring.NewRing(32, [576460752303439873])
    a, b = r.NewPoly(), r2.NewPoly()
    with pytest.raises(b200ring.B200Error):
        r.NTT(a, b)
    with pytest.raises(b200ring.B200Error):
        r.AtLevel(2)
    with pytest.raises(b200ring.B200Error):
        r.NTT(a, r.NewPoly(2))


# ---------------------------------------------------------------- BASELINE sizes (properties)
def test_config1_3n_roundtrip_degree_3x2_12():
    N, q = 3 * 2**12, 0x10000000000a1001
    r = b200ring.NewRing(N, [q])
    x = rand_poly(0xB2000001, 1, [q], N)
    p, y = r.NewPoly().set(x), r.NewPoly()
    r.NTT(p, y)
    s = O.ThreeNSubRing(N, q)
    assert (y.get()[0, 0] == s.NTT_horner(x[0, 0])).all()      # definitional O(N^2) check
    r.INTT(y, y)
    assert (y.get() == x).all()


def test_config3_headline_size_3n():
    """N = 3*2^15, L = 24: full parity on 2 polys x 24 limbs vs the oracle + Horner spot checks."""
    N, moduli, batch = 3 * 2**15, C3_PRIMES, 2
    r = b200ring.NewRing(N, moduli)
    host = rand_poly(0xB2000003, batch, moduli, N)
    p, o = r.NewPoly(batch).set(host), r.NewPoly(batch)
    r.NTT(p, o)
    got = o.get()
    E = O.totatives(3 * N)
    rng = np.random.default_rng(1)
    for l, q in enumerate(moduli):
        s = O.ThreeNSubRing(N, q)
        for b in range(batch):
            assert (got[b, l] == s.NTT(host[b, l])).all(), (b, l)
        if l % 8 == 0:
            for k in rng.integers(0, N, 8):
                assert s.horner_at(host[0, l], E[k]) == int(got[0, l, k])
    r.INTT(o, o)
    assert (o.get() == host).all()
    # linearity in the NTT domain: NTT(a + b) == NTT(a) + NTT(b)
    a, b2, sm = r.NewPoly(1).set(host[:1]), r.NewPoly(1).set(host[1:]), r.NewPoly(1)
    r.Add(a, b2, sm); r.NTT(sm, sm); r.NTT(a, a); r.NTT(b2, b2); r.Add(a, b2, a)
    assert (sm.get() == a.get()).all()


def test_config2_pow2_2_16_L24():
    N, moduli, batch = 1 << 16, QI60[:24], 2
    r = b200ring.NewRing(N, moduli)
    host = rand_poly(0xB2000002, batch, moduli, N)
    p, o = r.NewPoly(batch).set(host), r.NewPoly(batch)
    r.NTT(p, o)
    got = o.get()
    for l, q in enumerate(moduli):
        s = O.Pow2SubRing(N, q)
        for b in range(batch):
            assert (got[b, l] == s.NTT(host[b, l])).all(), (b, l)
    r.INTT(o, o)
    assert (o.get() == host).all()


def test_config4_polymul_3n_L16():
    """c = INTT(MulCoeffsMontgomery(MForm(NTT a), NTT b)) at N = 3*2^15, L = 16 vs the oracle; sparse
    operands vs schoolbook at N = 3*2^12."""
    N, moduli, batch = 3 * 2**15, C3_PRIMES[:16], 2
    r = b200ring.NewRing(N, moduli)
    ha, hb = rand_poly(0xB2000004, batch, moduli, N), rand_poly(0xB2000005, batch, moduli, N)
    a, b, c = r.NewPoly(batch).set(ha), r.NewPoly(batch).set(hb), r.NewPoly(batch)
    r.MulPoly(a, b, c)
    got = c.get()
    assert (a.get() == ha).all() and (b.get() == hb).all()      # operands untouched
    for l, q in enumerate(moduli):
        s = O.ThreeNSubRing(N, q)
        for bi in range(batch):
            fa, fb = s.NTT(ha[bi, l], order=1), s.NTT(hb[bi, l], order=1)
            prod = O.vec_op("MulCoeffsBarrett", q, fa, fb)
            assert (got[bi, l] == s.INTT(prod, order=1)).all(), (bi, l)
    N2, q2 = 3 * 2**12, 0x10000000000a1001
    r2 = b200ring.NewRing(N2, [q2])
    sa = np.zeros((1, 1, N2), dtype=np.uint64); sb = rand_poly(7, 1, [q2], N2)
    sa[0, 0, [0, 1, N2 // 2, N2 - 1]] = [3, q2 - 1, 5, 7]
    c2 = r2.NewPoly()
    r2.MulPoly(r2.NewPoly().set(sa), r2.NewPoly().set(sb), c2)
    assert (c2.get()[0, 0] == O.schoolbook_3n(N2, q2, sa[0, 0], sb[0, 0])).all()


def test_transform_host_pipelined_chunks():
    """b200_transform_host: HOST buffers streamed in chunks (H2D / kernels / D2H overlapped)."""
    rng = np.random.default_rng(11)
    for N, moduli, batch in ((1 << 16, QI60[:4], 150), (3 * 2**15, C3_PRIMES[:3], 70)):
        r = b200ring.NewRing(N, moduli)
        host = np.empty((batch, len(moduli), N), dtype=np.uint64)
        for l, q in enumerate(moduli):
            host[:, l, :] = rng.integers(0, q, (batch, N), dtype=np.uint64)
        out = np.empty_like(host)
        r.TransformHost(host, out, ["NTT"])
        for b in (0, batch // 2, batch - 1):            # first, middle and last chunk
            for l, q in enumerate(moduli):
                s = O.ThreeNSubRing(N, q) if N % 3 == 0 else O.Pow2SubRing(N, q)
                assert (out[b, l] == s.NTT(host[b, l])).all(), (N, b, l)
        back = np.empty_like(host)
        r.TransformHost(out, back, ["INTT"])
        assert (back == host).all()
        r.TransformHost(host, out, ["NTT", "INTT"])      # in one call, and in place
        assert (out == host).all()
        r.TransformHost(out, out, ["NTT", "INTT"])
        assert (out == host).all()


# power-of-two NTT-friendly primes the fold arithmetic admits (|c| <= 2^(k-33)): 2^60 + c and 2^55 + c (QI60 are 2^61 - c: Shoup path)
FOLD_POW2_PRIMES = [0x1000000000980001, 0x1000000000b00001, 0x80000000080001]


@pytest.mark.parametrize("env", [{"B200_GENERIC_COLS": "1"}, {"B200_NO_BIG": "1"}, {"B200_NO_FOLD": "1"}, {"B200_FOLD_COLS": "0"},
                                 {"B200_COLS_BAL": "0"}, {"B200_CW_BAL": "16"}])
@pytest.mark.parametrize("shape", [("std", 1 << 16), ("std", 1 << 15), ("m3", 3 * 2**15), ("m3", 3 * 2**14),
                                   ("std", 1 << 10), ("std", 1 << 11), ("m3", 3 * 2**10), ("m3", 3 * 2**11), ("m3", 3 * 2**16),
                                   ("stdf", 1 << 16), ("stdf", 1 << 13)])
def test_alternative_kernel_paths_agree(env, shape, monkeypatch):
    """The generic column-pass family, the shape-agnostic fallback (resident / layered kernels), the Shoup kernels
    (B200_NO_FOLD), the fold row pass behind Shoup columns (B200_FOLD_COLS=0) and the tiled / wide-tile fold column
    kernels must give the same words as the default path: the knobs are read at ring construction."""
    kind, N = shape
    if kind != "m3" and set(env) & {"B200_COLS_BAL", "B200_CW_BAL"}:
        pytest.skip("balanced column kernels exist for the 3N headline shape only")
    moduli = QI60[:2] if kind == "std" else FOLD_POW2_PRIMES if kind == "stdf" else O.find_3n_rns_primes(N, 60, 2)
    host = rand_poly(77, 2, moduli, N)
    ref = b200ring.NewRing(N, moduli)
    a, want = ref.NewPoly(2).set(host), ref.NewPoly(2)
    ref.NTT(a, want)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    alt = b200ring.NewRing(N, moduli)
    b, got = alt.NewPoly(2).set(host), alt.NewPoly(2)
    alt.NTT(b, got)
    assert (got.get() == want.get()).all()
    alt.INTT(got, got)
    assert (got.get() == host).all()
    sub = O.ThreeNSubRing(N, moduli[1]) if kind == "m3" else O.Pow2SubRing(N, moduli[1])
    assert (want.get()[1, 1] == sub.NTT(host[1, 1])).all()


@pytest.mark.parametrize("shape", [("std", 1 << 16), ("std", 1 << 13), ("m3", 3 * 2**15), ("m3", 3 * 2**14)])
def test_fused_product_on_fold_primes_equals_operator_sequence(shape):
    """b200_polymul on the fold path multiplies slot-wise with a Montgomery product and lets the inverse column pass give
    the 2^64 back; the operator sequence NTT, MForm, MulCoeffsMontgomery, INTT of
    ring/ring_test.go:106-110 runs through other kernels (no scaled constants) and must give the same words; a monomial
    operand pins the product itself (X^j * b is a signed rotation in the power-of-two ring)."""
    kind, N = shape
    moduli = FOLD_POW2_PRIMES[:2] if kind == "std" else C3_PRIMES[3:5] if N == 3 * 2**15 else O.find_3n_rns_primes(N, 60, 2)
    r = b200ring.NewRing(N, moduli)
    ha, hb = rand_poly(71, 2, moduli, N), rand_poly(72, 2, moduli, N)
    a, b, c, d, e = (r.NewPoly(2) for _ in range(5))
    a.set(ha); b.set(hb)
    r.MulPoly(a, b, c)
    r.NTT(a, d); r.NTT(b, e); r.MForm(d, d); r.MulCoeffsMontgomery(d, e, d); r.INTT(d, d)
    assert (c.get() == d.get()).all()
    if kind == "std":
        j = 5
        mono = np.zeros_like(ha); mono[:, :, j] = 1
        r.MulPoly(a.set(mono), b, c)
        got = c.get()
        for l, q in enumerate(moduli):
            want = np.concatenate([(np.uint64(q) - hb[:, l, N - j:]) % np.uint64(q), hb[:, l, : N - j]], axis=1)
            assert (got[:, l] == want).all()


def test_fold_arithmetic_extreme_inputs_headline_shape():
    """The fold kernels work on signed lazy words: all-(q-1), all-zero and alternating 0 / q-1 inputs drive the
    windows of tools/fold_ranges.py to their ends; results must still be the oracle's words (3N headline shape and the
    power-of-two shape on fold primes), forward and inverse."""
    for kind, N, moduli in (("m3", 3 * 2**15, C3_PRIMES[-2:]), ("std", 1 << 16, FOLD_POW2_PRIMES[:2])):
        r = b200ring.NewRing(N, moduli)
        subs = [O.ThreeNSubRing(N, q) if kind == "m3" else O.Pow2SubRing(N, q) for q in moduli]
        host = np.zeros((3, len(moduli), N), dtype=np.uint64)
        for l, q in enumerate(moduli):
            host[0, l, :] = q - 1
            host[1, l, ::2] = q - 1
            host[2, l, : N // 2] = q - 1
        p, o = r.NewPoly(3).set(host), r.NewPoly(3)
        r.NTT(p, o)
        got = o.get()
        for b in range(3):
            for l in range(len(moduli)):
                assert (got[b, l] == subs[l].NTT(host[b, l])).all(), (kind, b, l)
        r.INTT(o, o)
        assert (o.get() == host).all()
        # the inverse transform of the same extreme words taken as NTT-domain input
        o.set(host)
        r.INTT(o, o)
        back = o.get()
        for b in range(3):
            for l in range(len(moduli)):
                assert (back[b, l] == subs[l].INTT(host[b, l])).all(), (kind, b, l)


def test_3n_lazy_and_tree_order_flags_headline_shape():
    N, moduli = 3 * 2**15, C3_PRIMES[:2]
    r = b200ring.NewRing(N, moduli)
    host = rand_poly(5, 1, moduli, N)
    p, o, t = r.NewPoly().set(host), r.NewPoly(), r.NewPoly()
    r.NTT(p, o)
    canon = o.get()
    r.NTTLazy(p, o)
    lazy = o.get()
    for l, q in enumerate(moduli):
        assert (lazy[0, l] % np.uint64(q) == canon[0, l]).all() and int(lazy[0, l].max()) <= 6 * q - 2
    r.NTT(p, t, tree_order=True)
    s = O.ThreeNSubRing(N, moduli[0])
    assert (t.get()[0, 0] == s.NTT(host[0, 0], order=1)).all()
    assert (canon[0, 0] == t.get()[0, 0][s.go2fast()]).all()
    r.INTT(t, t, tree_order=True)
    assert (t.get() == host).all()
    r.INTTLazy(r.NewPoly().set(canon), o)
    lz = o.get()
    for l, q in enumerate(moduli):
        assert (lz[0, l] % np.uint64(q) == host[0, l]).all() and int(lz[0, l].max()) <= 2 * q - 1


@pytest.mark.parametrize("mform", [True, False])
def test_tensor_product_matches_reference_sequence(mform):
    """schemes/ckks/evaluator.go:821-834 (MForm first) and schemes/matrix_ckks/evaluator.go:163-173 (no MForm)."""
    N, moduli, batch = 3 * 2**10, O.find_3n_rns_primes(3 * 2**10, 60, 3), 2
    r = b200ring.NewRing(N, moduli)
    h = [rand_poly(90 + i, batch, moduli, N) for i in range(4)]
    a0, a1, b0, b1 = (r.NewPoly(batch).set(x) for x in h)
    c0, c1, c2 = r.NewPoly(batch), r.NewPoly(batch), r.NewPoly(batch)
    r.TensorProduct(a0, a1, b0, b1, c0, c1, c2, mform=mform)
    g0, g1, g2 = c0.get(), c1.get(), c2.get()
    for b in range(batch):
        for l, q in enumerate(moduli):
            x0, x1 = h[0][b, l], h[1][b, l]
            if mform:
                x0, x1 = O.vec_op("MForm", q, x0), O.vec_op("MForm", q, x1)
            w0 = O.vec_op("MulCoeffsMontgomery", q, x0, h[2][b, l])
            w2 = O.vec_op("MulCoeffsMontgomery", q, x1, h[3][b, l])
            w1 = O.vec_op("MulCoeffsMontgomery", q, x0, h[3][b, l])
            w1 = O.vec_op("MulCoeffsMontgomeryThenAdd", q, x1, h[2][b, l], w1)
            assert (g0[b, l] == w0).all() and (g1[b, l] == w1).all() and (g2[b, l] == w2).all(), (mform, b, l)
    # the same through the separate device operators, and the squaring case (c1 = 2 a0 a1)
    t0, t1, d0, d1, d2 = (r.NewPoly(batch) for _ in range(5))
    if mform:
        r.MForm(a0, t0); r.MForm(a1, t1)
    else:
        t0.set(h[0]); t1.set(h[1])
    r.MulCoeffsMontgomery(t0, b0, d0); r.MulCoeffsMontgomery(t1, b1, d2)
    r.MulCoeffsMontgomery(t0, b1, d1); r.MulCoeffsMontgomeryThenAdd(t1, b0, d1)
    assert (d0.get() == g0).all() and (d1.get() == g1).all() and (d2.get() == g2).all()
    r.TensorProduct(a0, a1, a0, a1, c0, c1, c2, mform=mform)
    r.MulCoeffsMontgomery(t0, a1, d1); r.Add(d1, d1, d1)
    assert (c1.get() == d1.get()).all()
    with pytest.raises(b200ring.B200Error):
        r.TensorProduct(a0, a1, b0, b1, c0, c0, c2)


@pytest.mark.parametrize("shape", [("std", 64, 3), ("std", 1 << 13, 4), ("m3", 24, 3), ("m3", 3 * 2**12, 3), ("m3", 3 * 2**15, 3)])
def test_rescale_div_by_last_modulus_ntt(shape):
    """Ring.DivRoundByLastModulusNTT / DivFloorByLastModulusNTT (ring/scaling.go:6-19,130-156) vs the oracle, and --
    like ring/ring_test.go:242-331 -- vs rounded big-integer division on a few coefficients."""
    kind, N, L = shape
    moduli = QI60[:L] if kind == "std" else (C3_PRIMES[:L] if N == 3 * 2**15 else O.find_3n_rns_primes(N, 60, L))
    rings = [O.Pow2SubRing(N, q) if kind == "std" else O.ThreeNSubRing(N, q) for q in moduli]
    r = b200ring.NewRing(N, moduli)
    batch = 2
    host = rand_poly(123, batch, moduli, N)              # NTT-domain words, canonical
    p0, p1 = r.NewPoly(batch).set(host), r.NewPoly(batch)
    for rnd in (True, False):
        (r.DivRoundByLastModulusNTT if rnd else r.DivFloorByLastModulusNTT)(p0, p1)
        got = p1.get()
        for b in range(batch):
            want = O.div_by_last_modulus_ntt(rings, host[b], round=rnd)
            assert (got[b, : L - 1] == want).all(), (shape, rnd, b)
    assert (p0.get() == host).all()
    # lower level through AtLevel, in place
    q = r.NewPoly(batch).set(host)
    r.AtLevel(L - 2).DivRoundByLastModulusNTT(q, q)
    got = q.get()
    want = O.div_by_last_modulus_ntt(rings[: L - 1], host[0, : L - 1], round=True)
    assert (got[0, : L - 2] == want).all() and (got[:, L - 2:] == host[:, L - 2:]).all()
    with pytest.raises(b200ring.B200Error):
        r.AtLevel(0).DivRoundByLastModulusNTT(p0, p1)
    # big-integer property on the first 4 coefficients (coefficient domain)
    if N <= 64:
        coeff = np.stack([rings[i].INTT(host[0, i]) for i in range(L)])
        r.DivRoundByLastModulusNTT(p0, p1)
        out = p1.get()[0]
        from test_oracle import _crt
        for j in range(4):
            x = _crt(coeff[:, j], moduli)
            w = (x + (moduli[-1] - 1) // 2) // moduli[-1]
            for i in range(L - 1):
                assert int(rings[i].INTT(out[i])[j]) == w % moduli[i]


def test_sub_then_mulscalar_two_modulus_op():
    N, moduli = 64, QI60[:2]
    r = b200ring.NewRing(N, moduli)
    h1, h2 = rand_poly(1, 1, moduli, N), rand_poly(2, 1, moduli, N)
    sc = [987654321987654321 % q for q in moduli]
    out = r.NewPoly()
    r.SubThenMulScalarMontgomeryTwoModulus(r.NewPoly().set(h1), r.NewPoly().set(h2), sc, out)
    for l, q in enumerate(moduli):
        assert (out.get()[0, l] == O.vec_op("SubThenMulScalarMontgomeryTwoModulus", q, h1[0, l], h2[0, l], scalar=sc[l])).all()


@pytest.mark.parametrize("N", [16, 24, 3 * 2**10, 1 << 12])
def test_remaining_vec_ops_bit_exact(N):
    """The rest of the element-wise family: SubRing.MulCoeffsLazy[ThenAddLazy], AddLazyThenMulScalarMontgomery, AddScalarLazy,
    AddScalarLazyThenNegTwoModulusLazy, AddScalarLazyThenMulScalarMontgomery, MulScalarMontgomeryThenAddScalar
    (ring/subring_ops.go:55-66,160-230), ZeroVec / MaskVec (ring/vec_ops.go:847-890), Ring.*DoubleRNSScalar* and
    MulByVectorMontgomery[ThenAddLazy] (ring/operations.go:165-266,366-377)."""
    is3n = N % 3 == 0
    moduli = O.find_3n_rns_primes(N, 60, 3) if is3n else QI60[18:21]
    r = b200ring.NewRing(N, moduli)
    batch, L = 2, len(moduli)
    h1, h2, h3 = (rand_poly(s, batch, moduli, N) for s in (41, 42, 43))
    h1[1] += h2[1]                                                   # lazy words in [0, 2q)
    p1, p2 = r.NewPoly(batch).set(h1), r.NewPoly(batch).set(h2)
    s0 = [1234567890123456789 % q for q in moduli]
    s1 = [987654321987654321 % q for q in moduli]

    def check(got, name, use2=False, use3=False, **kw):
        for b in range(batch):
            for l, q in enumerate(moduli):
                kws = {k: v[l] for k, v in kw.items()}
                want = O.vec_op(name, q, h1[b, l], h2[b, l] if use2 else None, h3[b, l] if use3 else None, **kws)
                assert (got[b, l] == want).all(), (name, b, l)

    for name, use3 in (("MulCoeffsLazy", False), ("MulCoeffsLazyThenAddLazy", True)):
        p3 = r.NewPoly(batch).set(h3)
        getattr(r, name)(p1, p2, p3)
        check(p3.get(), name, use2=True, use3=use3)
    p3 = r.NewPoly(batch).set(h3)
    r.AddLazyThenMulScalarMontgomery(p1, p2, s1, p3)
    check(p3.get(), "AddLazyThenMulScalarMontgomery", use2=True, scalar=s1)
    for name in ("AddScalarLazy", "AddScalarLazyThenNegTwoModulusLazy"):
        p3 = r.NewPoly(batch).set(h3)
        getattr(r, name)(p1, s0, p3)
        check(p3.get(), name, scalar=s0)
    for name in ("AddScalarLazyThenMulScalarMontgomery", "MulScalarMontgomeryThenAddScalar"):
        p3 = r.NewPoly(batch).set(h3)
        getattr(r, name)(p1, s0, s1, p3)
        check(p3.get(), name, scalar=s0, scalar1=s1)
    for w, mask in [(0, 0xffff), (13, (1 << 20) - 1), (63, 1), (64, 2**64 - 1)]:
        p3 = r.NewPoly(batch).set(h3)
        r.MaskVec(p1, w, mask, p3)
        check(p3.get(), "MaskVec", scalar=[w] * L, scalar1=[mask] * L)
    # Double operators take plain RNS scalars; the multiplications lift them to Montgomery form like operations.go:253-254
    hq = h1 % np.array(moduli, dtype=np.uint64)[None, :, None]
    pq = r.NewPoly(batch).set(hq)
    m0 = [v * 2**64 % q for v, q in zip(s0, moduli)]
    m1 = [v * 2**64 % q for v, q in zip(s1, moduli)]
    for name, oname, a0, a1, use3 in (("AddDoubleRNSScalar", "AddDoubleScalar", s0, s1, False), ("SubDoubleRNSScalar", "SubDoubleScalar", s0, s1, False),
                                      ("MulDoubleRNSScalar", "MulDoubleScalarMontgomery", m0, m1, False),
                                      ("MulDoubleRNSScalarThenAdd", "MulDoubleScalarMontgomeryThenAdd", m0, m1, True)):
        p3 = r.NewPoly(batch).set(h3)
        getattr(r, name)(pq, s0, s1, p3)
        got = p3.get()
        for b in range(batch):
            for l, q in enumerate(moduli):
                want = O.vec_op(oname, q, hq[b, l], None, h3[b, l] if use3 else None, scalar=a0[l], scalar1=a1[l])
                assert (got[b, l] == want).all(), (name, b, l)
    # one vector for every limb of every polynomial
    vec = r.NewPoly(1).set(h2[:1])
    for name, oname, use3 in (("MulByVectorMontgomery", "MulCoeffsMontgomery", False), ("MulByVectorMontgomeryThenAddLazy", "MulCoeffsMontgomeryThenAddLazy", True)):
        p3 = r.NewPoly(batch).set(h3)
        getattr(r, name)(p1, vec, p3)
        got = p3.get()
        for b in range(batch):
            for l, q in enumerate(moduli):
                want = O.vec_op(oname, q, h1[b, l], h2[0, 0], h3[b, l] if use3 else None)
                assert (got[b, l] == want).all(), (name, b, l)
    # ZeroVec at a lower level leaves the limbs above it alone
    p3 = r.NewPoly(batch).set(h3)
    r.AtLevel(L - 2).Zero(p3)
    got = p3.get()
    assert not got[:, : L - 1].any() and (got[:, L - 1] == h3[:, L - 1]).all()
    if N % 4:
        return
    with pytest.raises(b200ring.B200Error):
        r.AddScalarLazyThenMulScalarMontgomery(p1, s0[:1], s1, p3)


@pytest.mark.parametrize("shape", [("std", 64, 4), ("std", 1 << 12, 4), ("m3", 24, 4), ("m3", 3 * 2**12, 3)])
def test_div_by_last_modulus_coefficient_domain_and_many(shape):
    """Ring.Div{Floor,Round}ByLastModulus and ...Many[NTT] (ring/scaling.go:21-192) against the oracle's statement-by-statement
    restatement, batch of 2, in place and at a lower level."""
    kind, N, L = shape
    moduli = QI60[:L] if kind == "std" else O.find_3n_rns_primes(N, 60, L)
    rings = [O.Pow2SubRing(N, q) if kind == "std" else O.ThreeNSubRing(N, q) for q in moduli]
    r = b200ring.NewRing(N, moduli)
    batch = 2
    host = rand_poly(321, batch, moduli, N)
    p0 = r.NewPoly(batch).set(host)
    for rnd in (False, True):
        p1 = r.NewPoly(batch)
        (r.DivRoundByLastModulus if rnd else r.DivFloorByLastModulus)(p0, p1)
        got = p1.get()
        for b in range(batch):
            assert (got[b, : L - 1] == O.div_by_last_modulus(moduli, host[b], round=rnd)).all(), (shape, rnd, b)
        assert (p0.get() == host).all()
        for ntt in (False, True):
            name = "Div%sByLastModulusMany%s" % ("Round" if rnd else "Floor", "NTT" if ntt else "")
            for nb in range(L):
                p1 = r.NewPoly(batch).set(host[::-1].copy())
                getattr(r, name)(nb, p0, p1)
                got = p1.get()
                for b in range(batch):
                    want = O.div_by_last_modulus_many(rings, moduli, host[b], nb, round=rnd, ntt=ntt)
                    assert (got[b, : L - nb] == want).all(), (shape, name, nb, b)
                if nb:   # limbs above the result keep what they held
                    assert (got[:, L - 1] == host[::-1][:, L - 1]).all()
    # in place, at a lower level
    q = r.NewPoly(batch).set(host)
    r.AtLevel(L - 2).DivRoundByLastModulus(q, q)
    got = q.get()
    assert (got[1, : L - 2] == O.div_by_last_modulus(moduli[: L - 1], host[1, : L - 1], round=True)).all()
    assert (got[:, L - 2:] == host[:, L - 2:]).all()
    q = r.NewPoly(batch).set(host)
    r.DivFloorByLastModulusManyNTT(2, q, q)
    assert (q.get()[0, : L - 2] == O.div_by_last_modulus_many(rings, moduli, host[0], 2, round=False, ntt=True)).all()
    with pytest.raises(b200ring.B200Error):
        r.DivRoundByLastModulusMany(L, p0, p1)
    with pytest.raises(b200ring.B200Error):
        r.AtLevel(0).DivFloorByLastModulus(p0, p1)


def test_poly_copy_copylvl_zero_on_device():
    """Poly.Copy / CopyLvl / CopyNew / Zero (ring/poly.go:56-91) without a round trip through the host."""
    N, moduli = 64, QI60[:3]
    r = b200ring.NewRing(N, moduli)
    a, b = rand_poly(7, 2, moduli, N), rand_poly(8, 2, moduli, N)
    pa, pb = r.NewPoly(2).set(a), r.NewPoly(2).set(b)
    pb.CopyLvl(1, pa)
    got = pb.get()
    assert (got[:, :2] == a[:, :2]).all() and (got[:, 2] == b[:, 2]).all()
    pb.Copy(pa)
    assert pb.Equal(pa) and (pa.CopyNew().get() == a).all()
    pa.Copy(pa)
    assert (pa.get() == a).all()
    pb.Zero()
    assert not pb.get().any()
    r.AtLevel(0).NewPoly(2)
    with pytest.raises(b200ring.B200Error):
        pb.CopyLvl(3, pa)
    with pytest.raises(b200ring.B200Error):
        r.NewPoly(1).Copy(pa)


def test_fold_and_shoup_implementations_agree_on_structured_and_random_inputs():
    """tools/stress_fold_vs_shoup.py: the two independent transform programs (fold arithmetic / Shoup arithmetic, the latter
    forced with B200_NO_FOLD=1) on all 24 primes of BASELINE config 3 at N = 3*2^15, a generic 3N shape and two power-of-two
    rings on fold primes: all-(q-1), alternating, 2q-1 (top of the input contract) and random words through NTT, INTT,
    NTTLazy and the tree-order variants, bit for bit."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "stress_fold_vs_shoup.py"), "--rounds", "5"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count(": ok") == 4
