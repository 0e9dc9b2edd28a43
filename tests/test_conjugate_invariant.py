"""Conjugate-invariant rings Z_q[X + X^-1]/(X^2N + 1) (ring/ntt.go:80-124,716-786,1088-1158; NewRingConjugateInvariant,
ring/ring.go:280-282).  The reference pins this path with one property test, ring/ring_test.go:85-126: squaring through the
conjugate-invariant ring agrees with squaring the unfolded polynomial in the standard ring of degree 2N.  The standard path
is pinned bit-for-bit by the KATs of ring/ntt_test.go, so the property -- and the stronger one that the conjugate-invariant
NTT is the first half of the standard 2N-point NTT of the unfolded polynomial -- pins this one.  CPU tests check the oracle,
GPU tests compare the CUDA path with the oracle bit for bit through the C ABI."""
import json

import numpy as np
import pytest

from conftest import QI60
import b200ring
from oracle import oracle as O


def unfold_negated(p1, q, N):
    """ring_test.go:100-105: p2[j] = p1[j], p2[2N - j] = q - p1[j] (kept as written: a zero coefficient gives q)."""
    p2 = np.zeros(2 * N, dtype=np.uint64)
    p2[:N] = p1
    p2[2 * N - np.arange(1, N)] = np.uint64(q) - p1[1:]
    return p2


def square(sub, q, p):
    """NTT -> MForm -> MulCoeffsMontgomery(p, p) -> IMForm -> INTT (ring_test.go:107-119)."""
    a = sub.NTT(p)
    a = O.vec_op("MForm", q, a)
    a = O.vec_op("MulCoeffsMontgomery", q, a, a)
    a = O.vec_op("IMForm", q, a)
    return sub.INTT(a)


@pytest.mark.parametrize("N", [16, 32, 512, 4096])
def test_oracle_reference_property_and_first_half(N):
    for l, q in enumerate(QI60[:2]):
        ci, st = O.ConjugateInvariantSubRing(N, q), O.Pow2SubRing(2 * N, q)
        p1 = O.fill_uniform(0xC1, 0, l, q, N)
        p2 = unfold_negated(p1, q, N)
        assert (square(ci, q, p1) == square(st, q, p2)[:N]).all()            # ring_test.go:121-125
        assert (ci.NTT(p1) == st.NTT(p2)[:N]).all()                           # left half of the 2N-point transform
        assert (ci.INTT(ci.NTT(p1)) == p1).all()
        lazy = ci.NTT(p1, lazy=True)
        assert (lazy < np.uint64(6 * q)).all() and (lazy % np.uint64(q) == ci.NTT(p1)).all()
        # NthRoot = 4N: the tables are those of the standard ring of degree 2N, NInv = (2N)^-1 (subring.go:183-205)
        rf, rb = ci.tables()
        sf, sb = st.tables()
        assert (rf == sf).all() and (rb == sb).all() and ci.NInv == st.NInv and ci.PrimitiveRoot == st.PrimitiveRoot


def test_oracle_rejects_wrong_modulus():
    # 0x1fffffffffe00001 = 1 mod 2^21: valid up to 4N = 2^21
    assert O.ConjugateInvariantSubRing(1 << 19, QI60[0]).N == 1 << 19
    with pytest.raises(ValueError):
        O.ConjugateInvariantSubRing(1 << 20, QI60[0])


# ------------------------------------------------------------------------------------------- GPU
def rand_poly(seed, batch, moduli, N):
    out = np.empty((batch, len(moduli), N), dtype=np.uint64)
    for b in range(batch):
        for l, q in enumerate(moduli):
            out[b, l] = O.fill_uniform(seed, b, l, q, N)
    return out


@pytest.mark.gpu
def test_constructor_and_tables_bit_identical():
    r = b200ring.NewRingConjugateInvariant(1024, QI60[:2])
    assert r.Type() == b200ring.ConjugateInvariant and str(r.Type()) == "ConjugateInvariant" and r.NthRoot() == 4096
    assert b200ring.NewRingFromType(1024, QI60[:2], b200ring.ConjugateInvariant).NthRoot() == 4096
    for i, q in enumerate(r.ModuliChain()):
        s, sr = O.ConjugateInvariantSubRing(1024, q), r.SubRings[i]
        assert (sr.Modulus, sr.MRedConstant, sr.BRedConstant, sr.PrimitiveRoot, sr.NInv) == \
               (q, s.MRedConstant, s.BRedConstant, s.PrimitiveRoot, s.NInv)
        rf, rb = s.tables()
        assert len(sr.RootsForward) == 2048 and (sr.RootsForward == rf).all() and (sr.RootsBackward == rb).all()
    with pytest.raises(ValueError):                                          # q = 1 mod 2^21 only: 4N = 2^22 is refused
        b200ring.NewRingConjugateInvariant(1 << 20, QI60[:1])
    with pytest.raises(ValueError):
        b200ring.NewRingConjugateInvariant(3 << 10, QI60[:1])
    std = b200ring.NewRing(2048, QI60[:2])
    ci = std.ConjugateInvariantRing()
    assert ci.Type() == b200ring.ConjugateInvariant and ci.N() == 1024 and ci.ConjugateInvariantRing() is ci
    assert ci.StandardRing().N() == 2048 and std.StandardRing() is std


# NTT-friendly primes 2^60 + c and 2^55 + c (= 1 mod 2^18, |c| <= 2^(k-33)): the fold-arithmetic kernels run behind the folding layer
FOLD_PRIMES = [0x1000000000980001, 0x1000000000b00001, 0x80000000080001]


@pytest.mark.gpu
@pytest.mark.parametrize("logN,primes", [(n, "qi60") for n in (4, 6, 9, 10, 12, 13, 15, 16, 17)] + [(n, "fold") for n in (10, 13, 16)])
def test_ntt_matches_oracle(logN, primes):
    """Every kernel family behind the folding layer: resident (N <= 2^9 ... in shared memory), tuned two-pass
    (2^10..2^17) on Shoup arithmetic (Qi60) and on fold arithmetic (primes 2^k + c).  Forward, inverse, lazy variants, in place
    and out of place, bit for bit."""
    N = 1 << logN
    moduli = QI60[:3 if logN < 16 else 2] if primes == "qi60" else FOLD_PRIMES[:3 if logN < 16 else 2]
    batch = 3 if logN < 15 else 1
    r = b200ring.NewRingConjugateInvariant(N, moduli)
    host = rand_poly(0xC2 + logN, batch, moduli, N)
    subs = [O.ConjugateInvariantSubRing(N, q) for q in moduli]
    want = np.stack([np.stack([subs[l].NTT(host[b, l]) for l in range(len(moduli))]) for b in range(batch)])
    p, o = r.NewPoly(batch).set(host), r.NewPoly(batch)
    r.NTT(p, o)
    assert (o.get() == want).all()
    assert (p.get() == host).all()
    r.NTTLazy(p, o)
    lazy = o.get()
    for l, q in enumerate(moduli):
        assert (lazy[:, l] < np.uint64(6 * q)).all() and (lazy[:, l] % np.uint64(q) == want[:, l]).all()
    r.NTT(p, p)                                                               # in place
    assert (p.get() == want).all()
    r.INTT(p, o)
    assert (o.get() == host).all()
    r.INTTLazy(p, p)
    back = p.get()
    for l, q in enumerate(moduli):
        assert (back[:, l] < np.uint64(2 * q)).all() and (back[:, l] % np.uint64(q) == host[:, l]).all()
    # the inverse of arbitrary NTT-domain data (not an image of the forward transform of a canonical polynomial)
    y = rand_poly(0xC3 + logN, batch, moduli, N)
    wanti = np.stack([np.stack([subs[l].INTT(y[b, l]) for l in range(len(moduli))]) for b in range(batch)])
    p.set(y)
    r.INTT(p, o)
    assert (o.get() == wanti).all()


@pytest.mark.gpu
@pytest.mark.parametrize("N", [64, 4096])
def test_reference_property_through_the_c_abi(N):
    """ring/ring_test.go:85-126 with both rings on the GPU, plus Unfold and the fused product."""
    Q = QI60[:3]
    ring2n = b200ring.NewRing(2 * N, Q)
    ci = b200ring.NewRingFromType(N, Q, b200ring.ConjugateInvariant)
    p1 = rand_poly(0xC4, 2, Q, N)
    p2 = np.stack([np.stack([unfold_negated(p1[b, l], q, N) for l, q in enumerate(Q)]) for b in range(2)])
    a = ring2n.NewPoly(2).set(p2)
    ring2n.NTT(a, a); ring2n.MForm(a, a); ring2n.MulCoeffsMontgomery(a, a, a); ring2n.IMForm(a, a); ring2n.INTT(a, a)
    c = ci.NewPoly(2).set(p1)
    t = ci.NewPoly(2)
    ci.NTT(c, t); ci.MForm(t, t); ci.MulCoeffsMontgomery(t, t, t); ci.IMForm(t, t); ci.INTT(t, c)
    assert (c.get() == a.get()[:, :, :N]).all()
    # b200_polymul on the conjugate-invariant ring is the same product
    c2, out = ci.NewPoly(2).set(p1), ci.NewPoly(2)
    ci.MulPoly(c2, c2, out)
    assert (out.get() == c.get()).all()
    # UnfoldConjugateInvariantToStandard (ring/conjugate_invariant.go:7-23)
    u = ring2n.NewPoly(2)
    ci.UnfoldConjugateInvariantToStandard(c2, u)
    got = u.get()
    assert (got[:, :, :N] == p1).all() and (got[:, :, N:] == p1[:, :, ::-1]).all()
    with pytest.raises(b200ring.B200Error):
        ci.UnfoldConjugateInvariantToStandard(c2, ci.NewPoly(2))


@pytest.mark.gpu
def test_levels_host_buffers_and_literal():
    N, Q = 2048, QI60[:4]
    r = b200ring.NewRingConjugateInvariant(N, Q)
    host = rand_poly(0xC5, 2, Q, N)
    subs = [O.ConjugateInvariantSubRing(N, q) for q in Q]
    # AtLevel: only limbs 0..level are touched (ring/ntt.go:128)
    p = r.NewPoly(2).set(host)
    r.AtLevel(1).NTT(p, p)
    got = p.get()
    for l in range(4):
        for b in range(2):
            assert (got[b, l] == (subs[l].NTT(host[b, l]) if l <= 1 else host[b, l])).all()
    # host-buffer pipeline and the per-prime plugin granularity
    out = np.empty_like(host)
    r.TransformHost(host, out, ["NTT"])
    assert (out[:, 0] == np.stack([subs[0].NTT(host[b, 0]) for b in range(2)])).all()
    r.TransformHost(out, out, ["INTT"])
    assert (out == host).all()
    assert (r.SubRings[2].NTT(host[0, 2]) == subs[2].NTT(host[0, 2])).all()
    assert (r.SubRings[2].INTT(host[1, 2]) == subs[2].INTT(host[1, 2])).all()
    # ring literal: Type 1, NthRoot 4 (ring/subring.go:330-360), and back
    lit = json.loads(r.MarshalJSON())
    assert [(e["Type"], e["LogN"], e["NthRoot"]) for e in lit] == [(1, 11, 4)] * 4
    r2 = b200ring.UnmarshalJSON(r.MarshalJSON())
    assert r2.Type() == b200ring.ConjugateInvariant and r2.N() == N and r2.ModuliChain() == Q
    assert (r2.SubRings[3].RootsForward == r.SubRings[3].RootsForward).all()
