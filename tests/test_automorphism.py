"""ring/automorphism.go on the GPU: AutomorphismNTTIndex, AutomorphismNTT(+WithIndex, +ThenAddLazy), Automorphism, and
FoldStandardToConjugateInvariant (ring/conjugate_invariant.go:28-44).  The reference has no vectors for them; they are
pinned by definition (X -> X^gen evaluated directly) and by the commutation AutomorphismNTT(NTT p) == NTT(Automorphism p)
against the KAT-pinned transform, on the oracle (CPU) and bit for bit between oracle and CUDA path (GPU)."""
import numpy as np
import pytest

from conftest import QI60
import b200ring
from oracle import oracle as O


def direct_automorphism(N, q, gen, p):
    """X^i -> X^(i*gen) in Z_q[X]/(X^N + 1) with Python integers."""
    out = [0] * N
    for i in range(N):
        e = (i * gen) % (2 * N)
        v = int(p[i])
        if e >= N:
            e -= N
            v = (q - v) % q
        out[e] = (out[e] + v) % q
    return np.array(out, dtype=np.uint64)


@pytest.mark.parametrize("N", [16, 256])
def test_oracle_automorphism_definition_and_commutation(N):
    q = QI60[0]
    s = O.Pow2SubRing(N, q)
    p = O.fill_uniform(0xA0, 0, 0, q, N)
    for gen in (3, 5, 25, 2 * N - 1):
        got = O.automorphism(N, q, gen, p)
        assert (got % np.uint64(q) == direct_automorphism(N, q, gen, p)).all()
        idx = O.automorphism_ntt_index(N, 2 * N, gen)
        assert sorted(idx.tolist()) == list(range(N))
        assert (s.NTT(p)[idx.astype(np.int64)] == s.NTT(got % np.uint64(q))).all()
    ci = O.ConjugateInvariantSubRing(N, q)
    for gen in (5, 25):
        idx = O.automorphism_ntt_index(N, 4 * N, gen)
        assert (idx < N).all()
        assert (ci.NTT(p)[idx.astype(np.int64)] == ci.NTT(O.automorphism(N, q, gen, p, True) % np.uint64(q))).all()
    with pytest.raises(ValueError):
        O.automorphism_ntt_index(24, 48, 5)


def rand_poly(seed, batch, moduli, N):
    out = np.empty((batch, len(moduli), N), dtype=np.uint64)
    for b in range(batch):
        for l, q in enumerate(moduli):
            out[b, l] = O.fill_uniform(seed, b, l, q, N)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("N,typ", [(16, 0), (4096, 0), (1 << 16, 0), (16, 1), (2048, 1)])
def test_automorphisms_match_oracle(N, typ):
    Q = QI60[:3]
    ci = typ == 1
    r = b200ring.NewRingFromType(N, Q, b200ring.ConjugateInvariant if ci else b200ring.Standard)
    host = rand_poly(0xA1 + N, 2, Q, N)
    host[0, 0, 3] = 0                                                        # a negated zero is written as q (automorphism.go:147,171)
    p, o = r.NewPoly(2).set(host), r.NewPoly(2)
    for gen in ((5, 25, 125) if ci else (3, 5, 2 * N - 1, 5 ** 7 % (2 * N))):
        idx = O.automorphism_ntt_index(N, r.NthRoot(), gen)
        assert (b200ring.AutomorphismNTTIndex(N, r.NthRoot(), gen) == idx).all()
        want = host[:, :, idx.astype(np.int64)]
        r.AutomorphismNTT(p, gen, o)
        assert (o.get() == want).all()
        o.set(np.zeros_like(host))
        r.AutomorphismNTTWithIndex(p, idx, o)
        assert (o.get() == want).all()
        r.AutomorphismNTTWithIndexThenAddLazy(p, idx, o)                      # += , no reduction (:98-105)
        assert (o.get() == want + want).all()
        r.Automorphism(p, gen, o)
        wantc = np.stack([np.stack([O.automorphism(N, q, gen, host[b, l], ci) for l, q in enumerate(Q)]) for b in range(2)])
        assert (o.get() == wantc).all()
    # level: only limbs 0..level are written
    o.set(np.zeros_like(host))
    r.AtLevel(0).AutomorphismNTT(p, 5, o)
    assert (o.get()[:, 1:] == 0).all() and (o.get()[:, 0] != 0).any()
    # the reference's "cannot be in-place" and its panics
    with pytest.raises(b200ring.B200Error):
        r.AutomorphismNTT(p, 5, p)
    with pytest.raises(b200ring.B200Error):
        r.Automorphism(p, 5, p)
    with pytest.raises(b200ring.B200Error):
        r.AutomorphismNTTWithIndex(p, np.full(N, N, dtype=np.uint64), o)      # index out of range
    with pytest.raises(b200ring.B200Error):
        r.AutomorphismNTTWithIndex(p, np.zeros(N - 1, dtype=np.uint64), o)
    if ci:
        with pytest.raises(b200ring.B200Error):
            r.AutomorphismNTT(p, 3, o)                                        # slots of the other half


@pytest.mark.gpu
def test_automorphism_commutes_with_ntt_on_gpu_and_matrix_ring_refuses():
    N, Q = 1 << 13, QI60[:2]
    r = b200ring.NewRing(N, Q)
    host = rand_poly(0xA2, 1, Q, N)
    p, a, b = r.NewPoly().set(host), r.NewPoly(), r.NewPoly()
    gen = 5 ** 11 % (2 * N)
    r.NTT(p, a); r.AutomorphismNTT(a, gen, b)                                 # AutomorphismNTT(NTT p)
    r.Automorphism(p, gen, a); r.Reduce(a, a); r.NTT(a, a)                    # NTT(Automorphism p)
    assert (a.get() == b.get()).all()
    with pytest.raises(ValueError):
        b200ring.AutomorphismNTTIndex(24, 48, 5)
    m = b200ring.NewRing(24, O.find_3n_rns_primes(24, 60, 1))
    with pytest.raises(b200ring.B200Error):
        m.AutomorphismNTT(m.NewPoly(), 5, m.NewPoly())


@pytest.mark.gpu
def test_fold_standard_to_conjugate_invariant():
    """conjugate_invariant.go:28-44: ci[j] = st[index[j]] + st[j] mod q over the first N slots of a 2N polynomial."""
    N, Q = 1024, QI60[:2]
    st, ci = b200ring.NewRing(2 * N, Q), b200ring.NewRingConjugateInvariant(N, Q)
    host = rand_poly(0xA3, 2, Q, 2 * N)
    idx = O.automorphism_ntt_index(N, 4 * N, 4 * N - 1)                       # X -> X^-1 over the 2N-point slots: values < 2N
    assert (idx < 2 * N).all() and (idx >= N).any()
    ps, pc = st.NewPoly(2).set(host), ci.NewPoly(2)
    ci.FoldStandardToConjugateInvariant(ps, idx, pc)
    want = np.stack([np.stack([O.vec_op("Add", q, host[b, l, idx.astype(np.int64)], host[b, l, :N]) for l, q in enumerate(Q)])
                     for b in range(2)])
    assert (pc.get() == want).all()
    with pytest.raises(b200ring.B200Error):
        ci.FoldStandardToConjugateInvariant(pc, idx, pc)
