"""RNS basis extension, ModDown and the key-switch gadget product.

CPU part: the oracle restatement of ring/basis_extension.go against big-integer ground truth, as the
reference does in ring/ring_test.go:707-884 (testExtendBasis), and an exact end-to-end property of the
gadget product (a noiseless RNS gadget key returns (cx, cx*s) bit for bit).
GPU part: b200_mod_up / b200_mod_down_ntt / b200_gadget_product through the C ABI against that oracle."""
import numpy as np
import pytest

from conftest import QI60
import b200ring
from oracle import oracle as O


def chains(kind, N, LQ, LP):
    if kind == "std":
        Q, P = QI60[:LQ], QI60[LQ: LQ + LP]
        mk = lambda q: O.Pow2SubRing(N, q)
    else:
        pr = O.find_3n_rns_primes(N, 60, LQ + LP)
        Q, P = pr[:LQ], pr[LQ:]
        mk = lambda q: O.ThreeNSubRing(N, q)
    return Q, P, [mk(q) for q in Q], [mk(p) for p in P]


def gadget_key(Q, P, levelQ, levelP, N, s_ntt_Q, s_ntt_P):
    """Noiseless RNS gadget: g_d = P (Q/D_d) ((Q/D_d)^-1 mod D_d); component 0 = g_d, component 1 = g_d * s.
    NTT domain, Montgomery form (how rlwe stores evaluation keys, core/rlwe/keygenerator.go:70-71)."""
    nbPi, beta = levelP + 1, (levelQ + levelP + 1) // (levelP + 1)
    Qp, Pp = O._prod(Q[: levelQ + 1]), O._prod(P[: levelP + 1])
    key = []
    for d in range(beta):
        D = O._prod(Q[d * nbPi: min((d + 1) * nbPi, levelQ + 1)])
        g = Pp * (Qp // D) * pow(Qp // D, -1, D)
        comps = []
        for c in range(2):
            def limb(m, s):
                gm = (g % m) * pow(2, 64, m) % m
                return np.full(N, gm, dtype=np.uint64) if c == 0 else np.array([gm * int(v) % m for v in s], dtype=np.uint64)
            comps.append((np.stack([limb(q, s_ntt_Q[i]) for i, q in enumerate(Q[: levelQ + 1])]),
                          np.stack([limb(p, s_ntt_P[j]) for j, p in enumerate(P[: levelP + 1])])))
        key.append(comps)
    return key


def test_oracle_basis_extension_vs_big_integers():
    N = 16
    Q, P, rq, rp = chains("std", N, 4, 3)
    be = O.BasisExtender(Q, P)
    rng = np.random.default_rng(1)
    Qp, Pp = O._prod(Q), O._prod(P)
    coeffs = [int(rng.integers(0, 2**62)) ** 4 % Qp for _ in range(N)]
    polQ = np.array([[c % q for c in coeffs] for q in Q], dtype=np.uint64)
    polP = be.ModUpQtoP(3, 2, polQ)
    for j, p in enumerate(P):   # exact centred conversion
        assert [int(v) % p for v in polP[j]] == [(c if c <= Qp // 2 else c - Qp) % p for c in coeffs]
    back = be.ModUpPtoQ(2, 3, np.array([[c % p for c in coeffs] for p in P], dtype=np.uint64) % np.uint64(2**63))
    assert back.shape == (4, N)
    xs = [int(rng.integers(0, 2**62)) ** 7 % (Qp * Pp) for _ in range(N)]
    xQ = np.stack([rq[i].NTT(np.array([x % q for x in xs], dtype=np.uint64)) for i, q in enumerate(Q)])
    xP = np.stack([rp[j].NTT(np.array([x % p for x in xs], dtype=np.uint64)) for j, p in enumerate(P)])
    out = be.ModDownQPtoQNTT(3, 2, xQ, xP, rq, rp)
    for i, q in enumerate(Q):   # rounded division by P
        assert [int(v) for v in rq[i].INTT(out[i])] == [((x + Pp // 2) // Pp) % q for x in xs]
    # the coefficient-domain forms (:223-236, :264-279)
    cQ = np.array([[x % q for x in xs] for q in Q], dtype=np.uint64)
    cP = np.array([[x % p for x in xs] for p in P], dtype=np.uint64)
    out = be.ModDownQPtoQ(3, 2, cQ, cP)
    for i, q in enumerate(Q):
        assert [int(v) for v in out[i]] == [((x + Pp // 2) // Pp) % q for x in xs]
    out = be.ModDownQPtoP(3, 2, cQ, cP)
    centred = lambda x: x % Qp if x % Qp <= Qp // 2 else x % Qp - Qp
    for j, p in enumerate(P):   # (x - [x]_Q) / Q with the centred residue
        assert [int(v) for v in out[j]] == [((x - centred(x)) // Qp) % p for x in xs]


@pytest.mark.parametrize("kind,N", [("std", 16), ("m3", 24)])
@pytest.mark.parametrize("LQ,LP", [(5, 2), (4, 2), (3, 3), (6, 3)])
def test_oracle_gadget_product_exact_property(kind, N, LQ, LP):
    Q, P, rq, rp = chains(kind, N, LQ, LP)
    rng = np.random.default_rng(3)
    s = np.array(rng.integers(0, 3, N), dtype=np.uint64)
    sQ, sP = np.stack([r.NTT(s) for r in rq]), np.stack([r.NTT(s) for r in rp])
    key = gadget_key(Q, P, LQ - 1, LP - 1, N, sQ, sP)
    cx = np.stack([rng.integers(0, q, N, dtype=np.uint64) for q in Q])
    ct0, ct1 = O.gadget_product(rq, rp, LQ - 1, LP - 1, cx, key)
    assert (ct0 == cx).all()
    assert (ct1 == np.stack([O.vec_op("MulCoeffsBarrett", q, cx[i], sQ[i]) for i, q in enumerate(Q)])).all()


# ---------------------------------------------------------------------------------------------- GPU
def _qp_poly(ring, LQ, batch, q_part=None, p_part=None):
    L, N = ring.ModuliChainLength(), ring.N()
    h = np.zeros((batch, L, N), dtype=np.uint64)
    if q_part is not None:
        h[:, : q_part.shape[1]] = q_part
    if p_part is not None:
        h[:, LQ: LQ + p_part.shape[1]] = p_part
    return ring.NewPoly(batch).set(h)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,N,LQ,LP", [("std", 64, 4, 3), ("m3", 48, 5, 2), ("std", 1 << 12, 6, 2), ("std", 64, 20, 3), ("m3", 48, 18, 2)])
def test_mod_up_and_mod_down_bit_exact(kind, N, LQ, LP):
    # LQ = 18, 20: more than 16 source limbs in ModUpQtoP (the generic extension kernel; the accumulator's high word passes 2p there)
    Q, P, rq, rp = chains(kind, N, LQ, LP)
    ring = b200ring.NewRing(N, Q + P)
    ks = b200ring.KeySwitcher(ring, LQ)
    be = O.BasisExtender(Q, P)
    rng = np.random.default_rng(9)
    batch = 2
    hq = np.stack([np.stack([rng.integers(0, q, N, dtype=np.uint64) for q in Q]) for _ in range(batch)])
    hp = np.stack([np.stack([rng.integers(0, p, N, dtype=np.uint64) for p in P]) for _ in range(batch)])
    for levelQ, levelP in ((LQ - 1, LP - 1), (LQ - 2, LP - 1), (LQ - 1, 0)):
        pin, pout = _qp_poly(ring, LQ, batch, hq, hp), ring.NewPoly(batch)
        ks.ModUpQtoP(levelQ, levelP, pin, pout)
        got = pout.get()
        for b in range(batch):   # the lazy words of ModUpExact + SubScalarBigint, float64 correction included
            assert (got[b, LQ: LQ + levelP + 1] == be.ModUpQtoP(levelQ, levelP, hq[b])).all(), (levelQ, levelP, b)
        ks.ModUpPtoQ(levelP, levelQ, pin, pout)
        got = pout.get()
        for b in range(batch):
            assert (got[b, : levelQ + 1] == be.ModUpPtoQ(levelP, levelQ, hp[b])).all()
        ks.ModDownQPtoQNTT(levelQ, levelP, pin, pout)
        got = pout.get()
        for b in range(batch):
            want = be.ModDownQPtoQNTT(levelQ, levelP, hq[b][: levelQ + 1], hp[b][: levelP + 1], rq, rp)
            assert (got[b, : levelQ + 1] == want).all()
        ks.ModDownQPtoQ(levelQ, levelP, pin, pout)
        got = pout.get()
        for b in range(batch):
            assert (got[b, : levelQ + 1] == be.ModDownQPtoQ(levelQ, levelP, hq[b][: levelQ + 1], hp[b][: levelP + 1])).all()
        ks.ModDownQPtoP(levelQ, levelP, pin, pout)
        got = pout.get()
        for b in range(batch):
            assert (got[b, LQ: LQ + levelP + 1] == be.ModDownQPtoP(levelQ, levelP, hq[b][: levelQ + 1], hp[b][: levelP + 1])).all()
    with pytest.raises(b200ring.B200Error):
        ks.ModUpQtoP(LQ, 0, pin, pout)
    with pytest.raises(ValueError):
        b200ring.KeySwitcher(ring, LQ + LP)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,N,LQ,LP", [("std", 32, 5, 2), ("m3", 24, 4, 2), ("m3", 3 * 2**10, 6, 3), ("std", 1 << 13, 7, 2),
                                          ("m3", 3 * 2**15, 6, 2), ("m3", 24, 18, 17)])   # last: digits of 18 limbs, the canonical reduction of long digits
def test_gadget_product_matches_oracle_and_exact_property(kind, N, LQ, LP):
    Q, P, rq, rp = chains(kind, N, LQ, LP)
    ring = b200ring.NewRing(N, Q + P)
    ks = b200ring.KeySwitcher(ring, LQ)
    rng = np.random.default_rng(17)
    levelQ, levelP = LQ - 1, LP - 1
    beta = (levelQ + levelP + 1) // (levelP + 1)
    batch = 2
    cx = np.stack([np.stack([rng.integers(0, q, N, dtype=np.uint64) for q in Q]) for _ in range(batch)])
    # (1) random key: parity with the oracle restatement of rlwe.Evaluator.GadgetProduct
    key = [[(np.stack([rng.integers(0, q, N, dtype=np.uint64) for q in Q]), np.stack([rng.integers(0, p, N, dtype=np.uint64) for p in P]))
            for _ in range(2)] for _ in range(beta)]
    hk = np.zeros((2 * beta, LQ + LP, N), dtype=np.uint64)
    for d in range(beta):
        for c in range(2):
            hk[2 * d + c, :LQ], hk[2 * d + c, LQ:] = key[d][c]
    pkey = ring.NewPoly(2 * beta).set(hk)
    pcx, ct0, ct1 = _qp_poly(ring, LQ, batch, cx), ring.NewPoly(batch), ring.NewPoly(batch)
    ks.GadgetProduct(levelQ, levelP, pcx, pkey, ct0, ct1)
    g0, g1 = ct0.get(), ct1.get()
    if N <= 3 * 2**10:
        for b in range(batch):
            w0, w1 = O.gadget_product(rq, rp, levelQ, levelP, cx[b], key)
            assert (g0[b, :LQ] == w0).all() and (g1[b, :LQ] == w1).all(), (kind, N, b)
    # (2) noiseless gadget key: GadgetProduct(cx) == (cx, cx*s) exactly, at any size
    s = np.array(rng.integers(0, 3, N), dtype=np.uint64)
    sQ, sP = np.stack([r.NTT(s) for r in rq]), np.stack([r.NTT(s) for r in rp])
    gk = gadget_key(Q, P, levelQ, levelP, N, sQ, sP)
    for d in range(beta):
        for c in range(2):
            hk[2 * d + c, :LQ], hk[2 * d + c, LQ:] = gk[d][c]
    pkey.set(hk)
    ks.GadgetProduct(levelQ, levelP, pcx, pkey, ct0, ct1)
    g0, g1 = ct0.get(), ct1.get()
    assert (g0[:, :LQ] == cx).all()
    for i, q in enumerate(Q):
        for b in range(batch):
            assert (g1[b, i] == O.vec_op("MulCoeffsBarrett", q, cx[b, i], sQ[i])).all()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,N,LQ,LP", [("m3", 24, 4, 2), ("m3", 3 * 2**12, 6, 3), ("std", 1 << 12, 5, 2)])
def test_mul_relin_is_tensor_plus_gadget_product(kind, N, LQ, LP):
    """BASELINE config 5 (parity unpinned upstream): defined as the reference steps in sequence --
    tensor product (ckks.mulRelin), Relinearize = GadgetProduct(c2, rlk), Add.  With a noiseless key for s^2 the
    result is (c0 + c2*s2, c1) exactly."""
    Q, P, rq, rp = chains(kind, N, LQ, LP)
    ring = b200ring.NewRing(N, Q + P)
    ks = b200ring.KeySwitcher(ring, LQ)
    rng = np.random.default_rng(23)
    levelQ, levelP = LQ - 1, LP - 1
    beta = (levelQ + levelP + 1) // (levelP + 1)
    batch = 2
    h = [np.stack([np.stack([rng.integers(0, q, N, dtype=np.uint64) for q in Q]) for _ in range(batch)]) for _ in range(4)]
    a0, a1, b0, b1 = (_qp_poly(ring, LQ, batch, x) for x in h)
    s2 = np.array(rng.integers(0, 3, N), dtype=np.uint64)
    sQ, sP = np.stack([r.NTT(s2) for r in rq]), np.stack([r.NTT(s2) for r in rp])
    gk = gadget_key(Q, P, levelQ, levelP, N, sQ, sP)
    # rlk: component 0 carries g*s2, component 1 carries 0  (relinearisation key of a noiseless "ciphertext of s^2")
    hk = np.zeros((2 * beta, LQ + LP, N), dtype=np.uint64)
    for d in range(beta):
        hk[2 * d, :LQ], hk[2 * d, LQ:] = gk[d][1]
    rlk = ring.NewPoly(2 * beta).set(hk)
    o0, o1 = ring.NewPoly(batch), ring.NewPoly(batch)
    ks.MulRelin(levelQ, levelP, a0, a1, b0, b1, rlk, o0, o1, mform=True)
    # the same through the separate entry points
    c0, c1, c2, g0, g1 = (ring.NewPoly(batch) for _ in range(5))
    ring.AtLevel(levelQ).TensorProduct(a0, a1, b0, b1, c0, c1, c2, mform=True)
    ks.GadgetProduct(levelQ, levelP, c2, rlk, g0, g1)
    rl = ring.AtLevel(levelQ)
    rl.Add(c0, g0, g0); rl.Add(c1, g1, g1)
    assert (o0.get()[:, :LQ] == g0.get()[:, :LQ]).all() and (o1.get()[:, :LQ] == g1.get()[:, :LQ]).all()
    # rlwe.Evaluator.Relinearize as one call of the mirror (evaluator_evaluationkey.go:125-153) and ApplyEvaluationKey
    # (:110-117: out0 = c0 + gp0, out1 = gp1)
    r0, r1 = ring.NewPoly(batch), ring.NewPoly(batch)
    ks.Relinearize(levelQ, levelP, c0, c1, c2, rlk, r0, r1)
    assert (r0.get()[:, :LQ] == o0.get()[:, :LQ]).all() and (r1.get()[:, :LQ] == o1.get()[:, :LQ]).all()
    e0, e1, p0, p1 = (ring.NewPoly(batch) for _ in range(4))
    ks.ApplyEvaluationKey(levelQ, levelP, c0, c2, rlk, e0, e1)
    ks.GadgetProduct(levelQ, levelP, c2, rlk, p0, p1)
    rl.Add(c0, p0, p0)
    assert (e0.get()[:, :LQ] == p0.get()[:, :LQ]).all() and (e1.get()[:, :LQ] == p1.get()[:, :LQ]).all()
    # algebra: out0 = a0 b0 + (a1 b1) s2, out1 = a0 b1 + a1 b0   (slot-wise, mod q)
    got0, got1 = o0.get(), o1.get()
    for b in range(batch):
        for i, q in enumerate(Q):
            A0, A1, B0, B1 = ([int(v) for v in x[b, i]] for x in h)
            S = [int(v) for v in sQ[i]]
            assert [int(v) for v in got0[b, i]] == [(x * y + z * w % q * t) % q for x, y, z, w, t in zip(A0, B0, A1, B1, S)]
            assert [int(v) for v in got1[b, i]] == [(x * w + z * y) % q for x, y, z, w in zip(A0, B0, A1, B1)]


# ------------------------------------------------------------------ one special prime / power-of-two base
def pow2_gadget_key(Q, P, levelQ, pw2, sizes, N, s_ntt_Q, s_ntt_P):
    """Noiseless gadget for gadgetProductSinglePAndBitDecompLazy: g_{i,j} = P * 2^(j*pw2) * (Q/q_i) ((Q/q_i)^-1 mod q_i);
    component 0 = g, component 1 = g*s (NTT domain, Montgomery form).  sum_ij window_ij(cx_i) * g_ij = P*cx mod Q exactly."""
    Qp, Pp = O._prod(Q[: levelQ + 1]), P[0]
    key = []
    for i in range(levelQ + 1):
        qi = Q[i]
        base = Pp * (Qp // qi) * pow(Qp // qi, -1, qi)
        row = []
        for j in range(sizes[i]):
            g = base << (j * pw2)
            comps = []
            for c in range(2):
                def limb(m, s):
                    gm = (g % m) * pow(2, 64, m) % m
                    return np.full(N, gm, dtype=np.uint64) if c == 0 else np.array([gm * int(v) % m for v in s], dtype=np.uint64)
                comps.append((np.stack([limb(q, s_ntt_Q[u]) for u, q in enumerate(Q[: levelQ + 1])]),
                              np.stack([limb(P[0], s_ntt_P[0])])))
            row.append(comps)
        key.append(row)
    return key


@pytest.mark.parametrize("kind,N", [("std", 16), ("m3", 24)])
def test_oracle_single_special_prime_and_pow2_base(kind, N):
    LQ = 3
    Q, P, rq, rp = chains(kind, N, LQ, 1)
    rng = np.random.default_rng(5)
    s = np.array(rng.integers(0, 3, N), dtype=np.uint64)
    sQ, sP = np.stack([r.NTT(s) for r in rq]), np.stack([r.NTT(s) for r in rp])
    cx = np.stack([rng.integers(0, q, N, dtype=np.uint64) for q in Q])
    want1 = np.stack([O.vec_op("MulCoeffsBarrett", q, cx[i], sQ[i]) for i, q in enumerate(Q)])
    # mask == 0 (:213-215): one-limb digits
    ct0, ct1 = O.gadget_product(rq, rp, LQ - 1, 0, cx, gadget_key(Q, P, LQ - 1, 0, N, sQ, sP))
    assert (ct0 == cx).all() and (ct1 == want1).all()
    # power-of-two base: the windows of every limb rebuild cx exactly
    for pw2 in (16, 21):
        sizes = [(q.bit_length() + pw2 - 1) // pw2 for q in Q]              # rlwe/params.go:626-628
        ct0, ct1 = O.gadget_product_bit_decomp(rq, rp, LQ - 1, pw2, sizes, cx, pow2_gadget_key(Q, P, LQ - 1, pw2, sizes, N, sQ, sP))
        assert (ct0 == cx).all() and (ct1 == want1).all()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,N,LQ", [("std", 32, 3), ("m3", 24, 4), ("std", 1 << 12, 5), ("m3", 3 * 2**12, 3)])
def test_gadget_product_single_special_prime_and_pow2_base(kind, N, LQ):
    Q, P, rq, rp = chains(kind, N, LQ, 1)
    ring = b200ring.NewRing(N, Q + P)
    ks = b200ring.KeySwitcher(ring, LQ)
    rng = np.random.default_rng(29)
    batch = 2
    cx = np.stack([np.stack([rng.integers(0, q, N, dtype=np.uint64) for q in Q]) for _ in range(batch)])
    pcx, ct0, ct1 = _qp_poly(ring, LQ, batch, cx), ring.NewPoly(batch), ring.NewPoly(batch)
    s = np.array(rng.integers(0, 3, N), dtype=np.uint64)
    sQ, sP = np.stack([r.NTT(s) for r in rq]), np.stack([r.NTT(s) for r in rp])
    want1 = np.stack([np.stack([O.vec_op("MulCoeffsBarrett", q, cx[b, i], sQ[i]) for i, q in enumerate(Q)]) for b in range(batch)])
    for levelQ in (LQ - 1, LQ - 2):
        # (1) mask == 0 through b200_gadget_product with levelP == 0: random key against the oracle, then the exact property
        beta = levelQ + 1
        key = [[(np.stack([rng.integers(0, q, N, dtype=np.uint64) for q in Q]), np.stack([rng.integers(0, p, N, dtype=np.uint64) for p in P]))
                for _ in range(2)] for _ in range(beta)]
        hk = np.zeros((2 * beta, LQ + 1, N), dtype=np.uint64)
        for d in range(beta):
            for c in range(2):
                hk[2 * d + c, :LQ], hk[2 * d + c, LQ:] = key[d][c]
        pkey = ring.NewPoly(2 * beta).set(hk)
        ks.GadgetProduct(levelQ, 0, pcx, pkey, ct0, ct1)
        if N <= 1 << 12:
            for b in range(batch):
                w0, w1 = O.gadget_product(rq, rp, levelQ, 0, cx[b][: levelQ + 1], [[(kq[: levelQ + 1], kp) for kq, kp in d] for d in key])
                assert (ct0.get()[b, : levelQ + 1] == w0).all() and (ct1.get()[b, : levelQ + 1] == w1).all()
        # (2) power-of-two base
        for pw2 in (20, 13):
            sizes = [(q.bit_length() + pw2 - 1) // pw2 for q in Q[: levelQ + 1]]
            total = sum(sizes)
            rkey = [[[(np.stack([rng.integers(0, q, N, dtype=np.uint64) for q in Q[: levelQ + 1]]), np.stack([rng.integers(0, P[0], N, dtype=np.uint64)]))
                      for _ in range(2)] for _ in range(sizes[i])] for i in range(levelQ + 1)]
            gk = pow2_gadget_key(Q, P, levelQ, pw2, sizes, N, sQ, sP)
            for use, exact in ((rkey, False), (gk, True)):
                if not exact and N > 1 << 10:
                    continue                                                  # the oracle's per-window transforms are slow
                hk = np.zeros((2 * total, LQ + 1, N), dtype=np.uint64)
                k = 0
                for i in range(levelQ + 1):
                    for j in range(sizes[i]):
                        for c in range(2):
                            hk[k, : levelQ + 1], hk[k, LQ:] = use[i][j][c]
                            k += 1
                pk = ring.NewPoly(2 * total).set(hk)
                ks.GadgetProductBitDecomp(levelQ, 0, pw2, sizes, pcx, pk, ct0, ct1)
                g0, g1 = ct0.get(), ct1.get()
                if exact:
                    assert (g0[:, : levelQ + 1] == cx[:, : levelQ + 1]).all() and (g1[:, : levelQ + 1] == want1[:, : levelQ + 1]).all()
                else:
                    for b in range(batch):
                        w0, w1 = O.gadget_product_bit_decomp(rq[: levelQ + 1], rp, levelQ, pw2, sizes, cx[b][: levelQ + 1], use)
                        assert (g0[b, : levelQ + 1] == w0).all() and (g1[b, : levelQ + 1] == w1).all()
    with pytest.raises(b200ring.B200Error):
        ks.GadgetProductBitDecomp(LQ - 1, 0, 20, [1] * LQ, pcx, ring.NewPoly(1), ct0, ct1)      # key too small


@pytest.mark.gpu
@pytest.mark.parametrize("kind,N,LQ,LP", [("std", 32, 5, 2), ("m3", 48, 4, 1), ("m3", 3 * 2**10, 6, 3), ("std", 1 << 13, 7, 2)])
def test_hoisted_decomposition_and_gadget_product(kind, N, LQ, LP):
    """DecomposeNTT (evaluator_gadget_product.go:430-478) bit for bit against the oracle, and GadgetProductHoisted
    (:322-423) on it == GadgetProduct, for two keys sharing one decomposition and at a lower level."""
    Q, P, rq, rp = chains(kind, N, LQ, LP)
    ring = b200ring.NewRing(N, Q + P)
    ks = b200ring.KeySwitcher(ring, LQ)
    rng = np.random.default_rng(31)
    batch = 2
    cx = np.stack([np.stack([rng.integers(0, q, N, dtype=np.uint64) for q in Q]) for _ in range(batch)])
    pcx = _qp_poly(ring, LQ, batch, cx)
    for levelQ, levelP in ((LQ - 1, LP - 1), (LQ - 2, LP - 1)):
        beta = (levelQ + levelP + 1) // (levelP + 1)
        dec = ring.NewPoly(beta * batch)
        ks.DecomposeNTT(levelQ, levelP, pcx, dec)
        got = dec.get().reshape(beta, batch, LQ + LP, N)
        if N <= 3 * 2**10:
            for b in range(batch):
                want = O.decompose_ntt(rq, rp, levelQ, levelP, cx[b][: levelQ + 1])
                for d in range(beta):
                    assert (got[d, b, : levelQ + 1] == want[d][0]).all(), (d, b)
                    assert (got[d, b, LQ: LQ + levelP + 1] == want[d][1]).all(), (d, b)
        for seed in (1, 2):                                                   # two keys, one decomposition
            kr = np.random.default_rng(seed)
            hk = np.zeros((2 * beta, LQ + LP, N), dtype=np.uint64)
            for l, m in enumerate(Q + P):
                hk[:, l] = kr.integers(0, m, (2 * beta, N), dtype=np.uint64)
            pkey = ring.NewPoly(2 * beta).set(hk)
            h0, h1, g0, g1 = (ring.NewPoly(batch) for _ in range(4))
            ks.GadgetProductHoisted(levelQ, levelP, dec, pkey, h0, h1)
            ks.GadgetProduct(levelQ, levelP, pcx, pkey, g0, g1)
            assert (h0.get()[:, : levelQ + 1] == g0.get()[:, : levelQ + 1]).all()
            assert (h1.get()[:, : levelQ + 1] == g1.get()[:, : levelQ + 1]).all()
    with pytest.raises(b200ring.B200Error):
        ks.DecomposeNTT(LQ - 1, LP - 1, pcx, ring.NewPoly(1))


@pytest.mark.gpu
@pytest.mark.parametrize("N,LQ,LP", [(64, 4, 2), (1 << 12, 5, 1)])
def test_ciphertext_automorphism_is_keyswitch_add_permutation(N, LQ, LP):
    """rlwe.Evaluator.Automorphism / AutomorphismHoisted (core/rlwe/evaluator_automorphism.go:13-108): the same words as
    the reference's step sequence through the separate entry points, and the rotation identity with a noiseless key."""
    Q, P, rq, rp = chains("std", N, LQ, LP)
    ring = b200ring.NewRing(N, Q + P)
    ks = b200ring.KeySwitcher(ring, LQ)
    rng = np.random.default_rng(37)
    levelQ, levelP = LQ - 1, LP - 1
    beta = (levelQ + levelP + 1) // (levelP + 1)
    batch, gal = 2, 5 ** 3 % (2 * N)
    h0, h1 = (np.stack([np.stack([rng.integers(0, q, N, dtype=np.uint64) for q in Q]) for _ in range(batch)]) for _ in range(2))
    ct0, ct1 = _qp_poly(ring, LQ, batch, h0), _qp_poly(ring, LQ, batch, h1)
    s = np.array(rng.integers(0, 3, N), dtype=np.uint64)
    sQ, sP = np.stack([r.NTT(s) for r in rq]), np.stack([r.NTT(s) for r in rp])
    gk = gadget_key(Q, P, levelQ, levelP, N, sQ, sP)
    hk = np.zeros((2 * beta, LQ + LP, N), dtype=np.uint64)
    for d in range(beta):
        for c in range(2):
            hk[2 * d + c, :LQ], hk[2 * d + c, LQ:] = gk[d][c]
    key = ring.NewPoly(2 * beta).set(hk)
    o0, o1, g0, g1, a0, a1 = (ring.NewPoly(batch) for _ in range(6))
    ks.Automorphism(levelQ, levelP, ct0, ct1, gal, key, o0, o1)
    ks.GadgetProduct(levelQ, levelP, ct1, key, g0, g1)
    rl = ring.AtLevel(levelQ)
    rl.Add(g0, ct0, g0)
    rl.AutomorphismNTT(g0, gal, a0); rl.AutomorphismNTT(g1, gal, a1)
    assert (o0.get()[:, :LQ] == a0.get()[:, :LQ]).all() and (o1.get()[:, :LQ] == a1.get()[:, :LQ]).all()
    dec = ring.NewPoly(beta * batch)
    ks.DecomposeNTT(levelQ, levelP, ct1, dec)
    ks.Automorphism(levelQ, levelP, ct0, ct1, gal, key, g0, g1, decomp=dec)
    assert (g0.get()[:, :LQ] == o0.get()[:, :LQ]).all() and (g1.get()[:, :LQ] == o1.get()[:, :LQ]).all()
    # noiseless key (g, g*s): GadgetProduct(ct1) = (ct1, ct1*s), so out = phi(ct0 + ct1), phi(ct1 * s) slot-wise
    idx = O.automorphism_ntt_index(N, 2 * N, gal).astype(np.int64)
    got0, got1 = o0.get(), o1.get()
    for i, q in enumerate(Q):
        for b in range(batch):
            assert (got0[b, i] == O.vec_op("Add", q, h0[b, i], h1[b, i])[idx]).all()
            assert (got1[b, i] == O.vec_op("MulCoeffsBarrett", q, h1[b, i], sQ[i])[idx]).all()
    m = b200ring.NewRing(24, O.find_3n_rns_primes(24, 60, 3))
    with pytest.raises(b200ring.B200Error):                                   # 3N rings have no automorphism tables (automorphism.go:14-16)
        km = b200ring.KeySwitcher(m, 2)
        km.Automorphism(1, 0, m.NewPoly(), m.NewPoly(), 5, m.NewPoly(4), m.NewPoly(), m.NewPoly())
