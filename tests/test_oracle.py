"""CPU-only tests that pin the oracle (oracle/ref_ring.c) before it is trusted as the checker.

* power-of-two path: the 6 known-answer vectors of ring/ntt_test.go:10-89 (tests/golden/ntt_kat.json)
* modular primitives: the property test of ring/ring_test.go:534-670 (BRed/MRed/MForm vs big ints
  on the edge values 1, q-1, 2^64-1 for the 14 primes Qi60[18:32] of ring/test_params.go)
* 3N path: vectors from the reference's Python prototype (tests/golden/ntt3n_proto.json), the
  definition ring/ntt_3n.go:99-108 (Horner / Gaussian), round trips and schoolbook products as in
  ring/ntt_3n_test.go:79-222,367-554.
"""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, QI60
from oracle import oracle as O


def test_kat_pow2_forward_inverse_and_tables():
    kat = json.load(open(os.path.join(GOLDEN, "ntt_kat.json")))["vectors"]
    assert [v["N"] for v in kat] == [16, 32, 64, 128, 256, 512]
    for v in kat:
        for l, q in enumerate(v["Qis"]):
            s = O.Pow2SubRing(v["N"], q)
            x = np.array(v["poly"][l], dtype=np.uint64)
            want = np.array(v["polyNTT"][l], dtype=np.uint64)
            got = s.NTT(x)
            assert (got == want).all(), (v["N"], l)
            assert (s.INTT(want) == x).all()
            lazy = s.NTT(x, lazy=True)
            assert (lazy % np.uint64(q) == want).all() and int(lazy.max()) <= 6 * q - 2
            if v["N"] <= 64:  # definitional: out[i] = f(psi^(2 brev(i)+1))
                assert O.py_pow2_ntt(v["N"], q, s.PrimitiveRoot, v["poly"][l]) == [int(t) for t in got]
            # table invariants of subring.go:187-205
            rf, rb = s.tables()
            one = O.lib().ref_mform(1, q, O._p(O.bred_constant(q)))
            assert int(rf[0]) == one == int(rb[0])
            prod = [O.lib().ref_mred(int(a), int(b), q, s.MRedConstant) for a, b in zip(rf, rb)]
            assert all(p == one for p in prod)  # psi^k * psi^-k = 1 (Montgomery form)


def test_modular_reduction_edge_values():
    L = O.lib()
    for q in QI60[18:32]:
        brc = O.bred_constant(q)
        qinv = L.ref_gen_mred_constant(q)
        assert (qinv * q) % 2**64 == 1
        assert (int(brc[0]) << 64 | int(brc[1])) == 2**128 // q
        for x in (1, q - 1, 2**64 - 1):
            assert L.ref_bred_add(x, q, O._p(brc)) == x % q
            assert L.ref_mform(x, q, O._p(brc)) == (x << 64) % q
            assert L.ref_imform(x, q, qinv) == (x * pow(2**64, -1, q)) % q
            for y in (1, q - 1, 2**64 - 1):
                assert L.ref_bred(x, y, q, O._p(brc)) == (x * y) % q
                ym = L.ref_mform(y, q, O._p(brc))
                assert L.ref_mred(x, ym, q, qinv) == (x * y) % q
                lazy = L.ref_mred_lazy(x % q, ym, q, qinv)
                assert lazy % q == (x * y) % q and lazy < 2 * q


def test_vec_ops_against_python_ints():
    q = QI60[0]
    rng = np.random.default_rng(7)
    N = 64
    a = rng.integers(0, q, N, dtype=np.uint64)
    b = rng.integers(0, q, N, dtype=np.uint64)
    c = rng.integers(0, q, N, dtype=np.uint64)
    rinv = pow(2**64, -1, q)
    A, B, Cc = [int(x) for x in a], [int(x) for x in b], [int(x) for x in c]
    chk = lambda got, want: [int(g) for g in got] == want
    assert chk(O.vec_op("Add", q, a, b), [(x + y) % q for x, y in zip(A, B)])
    assert chk(O.vec_op("Sub", q, a, b), [(x - y) % q for x, y in zip(A, B)])
    assert chk(O.vec_op("MulCoeffsBarrett", q, a, b), [(x * y) % q for x, y in zip(A, B)])
    assert chk(O.vec_op("MulCoeffsMontgomery", q, a, b), [(x * y * rinv) % q for x, y in zip(A, B)])
    assert chk(O.vec_op("MulCoeffsMontgomeryThenAdd", q, a, b, c), [(z + x * y * rinv) % q for x, y, z in zip(A, B, Cc)])
    assert chk(O.vec_op("MulCoeffsMontgomeryThenSub", q, a, b, c), [(z - x * y * rinv) % q for x, y, z in zip(A, B, Cc)])
    assert chk(O.vec_op("MForm", q, a), [(x << 64) % q for x in A])
    assert chk(O.vec_op("IMForm", q, a), [(x * rinv) % q for x in A])
    big = rng.integers(0, 2**63, N, dtype=np.uint64) * np.uint64(2) + np.uint64(1)
    assert chk(O.vec_op("Reduce", q, big), [int(x) % q for x in big])


def test_remaining_vec_ops_against_python_ints():
    """The rest of ring/vec_ops.go / ring/subring_ops.go:55-66,160-230 and the Double operators of ring/operations.go:165-266,
    word for word against Python integers (lazy results are the Go expressions modulo 2^64)."""
    q = QI60[3]
    rng = np.random.default_rng(11)
    N, M = 64, 1 << 64
    a = rng.integers(0, 2 * q, N, dtype=np.uint64)          # lazy inputs
    b = rng.integers(0, 2 * q, N, dtype=np.uint64)
    c = rng.integers(0, q, N, dtype=np.uint64)
    rinv = pow(2**64, -1, q)
    A, B, Cc = [int(x) for x in a], [int(x) for x in b], [int(x) for x in c]
    s0, s1 = 1234567890123456789 % q, 987654321987654321 % q
    ints = lambda got: [int(g) for g in got]
    assert ints(O.vec_op("ZeroVec", q, a, None, c)) == [0] * N
    assert ints(O.vec_op("MulCoeffsLazy", q, a, b)) == [x * y % M for x, y in zip(A, B)]
    assert ints(O.vec_op("MulCoeffsLazyThenAddLazy", q, a, b, c)) == [(z + x * y) % M for x, y, z in zip(A, B, Cc)]
    assert ints(O.vec_op("AddLazyThenMulScalarMontgomery", q, a, b, scalar=s1)) == [(x + y) * s1 * rinv % q for x, y in zip(A, B)]
    assert ints(O.vec_op("AddScalarLazy", q, a, scalar=s0)) == [(x + s0) % M for x in A]
    assert ints(O.vec_op("AddScalarLazyThenNegTwoModulusLazy", q, a, scalar=s0)) == [(s0 + 2 * q - x) % M for x in A]
    assert ints(O.vec_op("AddScalarLazyThenMulScalarMontgomery", q, a, scalar=s0, scalar1=s1)) == [(x + s0) * s1 * rinv % q for x in A]
    assert ints(O.vec_op("MulScalarMontgomeryThenAddScalar", q, a, scalar=s0, scalar1=s1)) == [(x * s1 * rinv + s0) % q for x in A]
    for w, mask in [(0, 0xffff), (13, (1 << 20) - 1), (63, 1), (64, M - 1), (200, M - 1)]:
        assert ints(O.vec_op("MaskVec", q, a, scalar=w, scalar1=mask)) == [(x >> w) & mask for x in A]
    h = N // 2
    Cq = [x % q for x in A]
    cq = np.array(Cq, dtype=np.uint64)
    assert ints(O.vec_op("AddDoubleScalar", q, cq, scalar=s0, scalar1=s1)) == [(x + (s0 if i < h else s1)) % q for i, x in enumerate(Cq)]
    assert ints(O.vec_op("SubDoubleScalar", q, cq, scalar=s0, scalar1=s1)) == [(x - (s0 if i < h else s1)) % q for i, x in enumerate(Cq)]
    assert ints(O.vec_op("MulDoubleScalarMontgomery", q, cq, scalar=s0, scalar1=s1)) == [x * (s0 if i < h else s1) * rinv % q for i, x in enumerate(Cq)]
    assert ints(O.vec_op("MulDoubleScalarMontgomeryThenAdd", q, cq, None, c, scalar=s0, scalar1=s1)) == [
        (z + x * (s0 if i < h else s1) * rinv) % q for i, (x, z) in enumerate(zip(Cq, Cc))]


def test_3n_prototype_vectors():
    vec = json.load(open(os.path.join(GOLDEN, "ntt3n_proto.json")))["vectors"]
    assert len(vec) >= 8
    for v in vec:
        N, p, w = v["N"], v["p"], v["w"]
        r = O.ThreeNSubRing(N, p, w)
        assert [int(z) for z in r.zetas()] == v["zetas"]
        assert [int(e) for e in r.leaf_exponents()] == v["leaf_exponents"]
        x = np.array(v["coeffs"], dtype=np.uint64)
        tree = r.NTT(x, order=1)
        assert [int(t) for t in tree] == v["ntt_tree_order"]
        assert (r.INTT(tree, order=1) == x).all()
        go = r.NTT(x)  # Go order: slot k <-> w^E[k], E ascending totatives (ntt_3n.go:46-53)
        assert [int(t) for t in go] == O.py_3n_ntt(N, p, w, v["coeffs"])
        assert (go[np.arange(N)] == tree[r.go2fast()]).all()


@pytest.mark.parametrize("N", [6, 12, 18, 24, 36, 48, 96, 192])
def test_3n_definition_roundtrip_product(N):
    q = O.find_3n_rns_primes(N, 60, 1)[0]
    r = O.ThreeNSubRing(N, q)
    assert pow(r.omega, 3 * N, q) == 1 and pow(r.omega, 3 * N // 2, q) != 1 and pow(r.omega, N, q) != 1
    x = O.fill_uniform(11, 0, 0, q, N)
    y = r.NTT(x)
    assert (r.NTT_horner(x) == y).all()                     # ntt_3n.go:99-108
    assert [int(t) for t in y] == O.py_3n_ntt(N, q, r.omega, x)
    assert (r.INTT(y) == x).all()                           # ntt_3n_test.go:139-222
    if N <= 96:
        assert (r.INTT_gauss(y) == x).all()                 # ntt_3n.go:118-151,170-222
    z = x.copy()
    for _ in range(8):                                      # 8 repeated cycles, ntt_3n_test.go:183-199
        z = r.INTT(r.NTT(z))
    assert (z == x).all()
    zero = np.zeros(N, dtype=np.uint64)
    one = zero.copy(); one[0] = 1
    assert (r.NTT(zero) == 0).all() and (r.NTT(one) == 1).all()
    a = O.fill_uniform(12, 0, 0, q, N)
    b = O.fill_uniform(13, 0, 0, q, N)
    prod = np.array([(int(u) * int(v)) % q for u, v in zip(r.NTT(a), r.NTT(b))], dtype=np.uint64)
    assert (r.INTT(prod) == O.schoolbook_3n(N, q, a, b)).all()   # ntt_3n_test.go:455-554


def test_3n_fixed_product_n12():
    # fixed operands of ring/ntt_3n_test.go:79-135 style: schoolbook mod X^12 - X^6 + 1, by hand
    N, q = 12, O.find_3n_rns_primes(12, 20, 1)[0]
    a = np.zeros(N, dtype=np.uint64); b = np.zeros(N, dtype=np.uint64)
    a[11] = 1; b[1] = 1                      # X^11 * X = X^12 = X^6 - 1
    c = O.schoolbook_3n(N, q, a, b)
    want = np.zeros(N, dtype=np.uint64); want[6] = 1; want[0] = q - 1
    assert (c == want).all()
    r = O.ThreeNSubRing(N, q)
    prod = np.array([(int(u) * int(v)) % q for u, v in zip(r.NTT(a), r.NTT(b))], dtype=np.uint64)
    assert (r.INTT(prod) == want).all()


def test_3n_config1_degree_3x2_12():
    # BASELINE config 1: forward+inverse round trip, degree 3*2^12, single 3N-friendly prime
    N = 3 * 2**12
    q = O.find_3n_rns_primes(N, 60, 1)[0]
    assert q == 0x10000000000a1001
    r = O.ThreeNSubRing(N, q)
    x = O.fill_uniform(0xB2000001, 0, 0, q, N)
    y = r.NTT(x)
    assert (r.NTT_horner(x) == y).all()
    assert (r.INTT(y) == x).all()


def test_3n_headline_size_spot_checks():
    N = 3 * 2**15
    qs = O.find_3n_rns_primes(N, 60, 24)
    assert qs[0] == 0x1000000000248001 and qs[-1] == 0x10000000042b0001
    r = O.ThreeNSubRing(N, qs[5])
    x = O.fill_uniform(0xB2000003, 0, 5, qs[5], N)
    y = r.NTT(x)
    E = O.totatives(3 * N)
    assert len(E) == N
    rng = np.random.default_rng(3)
    for k in list(rng.integers(0, N, 48)) + [0, 1, N - 1]:
        assert r.horner_at(x, E[k]) == int(y[k])
    assert (r.INTT(y) == x).all()


def test_primes_and_roots():
    assert O.lib().ref_primitive_root(576460752303439873) >= 3
    for q in QI60[:4]:
        assert O.lib().ref_is_prime(q) and not O.lib().ref_is_prime(q + 2)
        f = np.zeros(16, dtype=np.uint64)
        n = O.lib().ref_unique_factors(q - 1, O._p(f))
        rem = q - 1
        for p in f[:n]:
            p = int(p); assert O.lib().ref_is_prime(p)
            while rem % p == 0:
                rem //= p
        assert rem == 1


def _crt(residues, moduli):
    x, M = 0, 1
    for r, q in zip(residues, moduli):
        t = ((int(r) - x) * pow(M, -1, q)) % q
        x += M * t
        M *= q
    return x


@pytest.mark.parametrize("kind", ["std", "m3"])
def test_rescale_matches_big_integer_division(kind):
    """ring/ring_test.go:242-331 (testDiv{Floor,Round}ByLastModulusMany): the restated DivRound/DivFloorByLastModulusNTT
    equals rounded / floored division of the CRT-reconstructed coefficients by the last modulus."""
    N = 32 if kind == "std" else 24
    moduli = QI60[:3] if kind == "std" else O.find_3n_rns_primes(N, 60, 3)
    rings = [O.Pow2SubRing(N, q) if kind == "std" else O.ThreeNSubRing(N, q) for q in moduli]
    rng = np.random.default_rng(5)
    Q = 1
    for q in moduli:
        Q *= q
    coeffs = [int(rng.integers(0, 2**62)) * int(rng.integers(0, 2**62)) * int(rng.integers(0, 2**50)) % (Q // 10) for _ in range(N)]
    p0 = np.array([[c % q for c in coeffs] for q in moduli], dtype=np.uint64)
    p0_ntt = np.stack([r.NTT(p0[i]) for i, r in enumerate(rings)])
    for rnd in (True, False):
        got = O.div_by_last_modulus_ntt(rings, p0_ntt, round=rnd)
        ql = moduli[-1]
        want = [((c + (ql - 1) // 2) // ql) if rnd else (c // ql) for c in coeffs]
        for i in range(len(moduli) - 1):
            back = rings[i].INTT(got[i])
            assert [int(v) for v in back] == [w % moduli[i] for w in want], (kind, rnd, i)


@pytest.mark.parametrize("kind", ["std", "m3"])
def test_div_by_last_modulus_many_matches_big_integer_division(kind):
    """ring/ring_test.go:242-331 (testDivFloorByLastModulusMany / testDivRoundByLastModulusMany): level successive
    divisions of coefficients below Q/10 equal floored / rounded big-integer division, for the coefficient-domain and the
    NTT-domain forms of ring/scaling.go:21-192."""
    N = 32 if kind == "std" else 24
    L = 4
    moduli = QI60[:L] if kind == "std" else O.find_3n_rns_primes(N, 60, L)
    rings = [O.Pow2SubRing(N, q) if kind == "std" else O.ThreeNSubRing(N, q) for q in moduli]
    rng = np.random.default_rng(9)
    Q = 1
    for q in moduli:
        Q *= q
    coeffs = [int.from_bytes(rng.bytes(32), "little") % (Q // 10) for _ in range(N)]
    p0 = np.array([[c % q for c in coeffs] for q in moduli], dtype=np.uint64)
    p0_ntt = np.stack([r.NTT(p0[i]) for i, r in enumerate(rings)])
    for rnd in (False, True):
        for nb in range(L):
            want = list(coeffs)
            for j in range(nb):
                ql = moduli[L - 1 - j]
                want = [(c + (ql - 1) // 2) // ql if rnd else c // ql for c in want]
            got = O.div_by_last_modulus_many(rings, moduli, p0, nb, round=rnd, ntt=False)
            assert got.shape[0] == L - nb
            for i in range(L - nb):
                assert [int(v) for v in got[i]] == [w % moduli[i] for w in want], (kind, rnd, nb, i)
            got = O.div_by_last_modulus_many(rings, moduli, p0_ntt, nb, round=rnd, ntt=True)
            for i in range(L - nb):
                assert [int(v) for v in rings[i].INTT(got[i])] == [w % moduli[i] for w in want], (kind, rnd, nb, i, "ntt")


def test_3n_small_prime_cases_of_the_reference():
    """ring/ntt_3n_test.go:139-222 with the reference's own parameters (31-bit primes from Find3NRNSPrimes(N, 31, 1, 1<<16),
    N = 12 and 24): round trip, eight Forward+Backward cycles, the zero and the all-ones vector; plus the definition
    out[k] = f(omega^E[k]) (ntt_3n.go:99-108) for these small moduli."""
    rng = np.random.default_rng(31)
    for N in (12, 24):
        q = O.find_3n_rns_primes(N, 31, 1)[0]
        assert q.bit_length() == 32 and q % (3 * N) == 1
        s = O.ThreeNSubRing(N, q)
        ref = rng.integers(0, q, N, dtype=np.uint64)
        cur = ref.copy()
        for _ in range(8):
            cur = s.INTT(s.NTT(cur))
        assert (cur == ref).all()
        for v in (np.zeros(N, dtype=np.uint64), np.ones(N, dtype=np.uint64)):
            assert (s.INTT(s.NTT(v)) == v).all()
        assert not s.NTT(np.zeros(N, dtype=np.uint64)).any()
        E = [int(e) for e in O.totatives(3 * N)]
        got = s.NTT(ref)
        for k in range(N):
            x = pow(s.omega, E[k], q)
            assert int(got[k]) == sum(int(c) * pow(x, j, q) for j, c in enumerate(ref)) % q


def test_modular_reduction_random_words_property():
    """ring/ring_test.go:534-670 (testModularReduction / testMForm) widened from the three edge values to seeded random 64-bit
    words: every primitive of ring/modular_reduction.go equals big-integer arithmetic, for 61-bit, 60-bit 3N and 32-bit moduli."""
    L = O.lib()
    rng = np.random.default_rng(2024)
    moduli = list(QI60[:4]) + list(O.find_3n_rns_primes(98304, 60, 2)) + list(O.find_3n_rns_primes(24, 31, 1)) + [0xffffffffffffffc5 >> 3 | 1]
    moduli = [q for q in moduli if q % 2 == 1]
    M = 1 << 64
    for q in moduli[:7]:
        brc = O.bred_constant(q)
        qinv = L.ref_gen_mred_constant(q)
        rinv = pow(M, -1, q)
        xs = [int(v) for v in rng.integers(0, M, 200, dtype=np.uint64)]
        ys = [int(v) for v in rng.integers(0, M, 200, dtype=np.uint64)]
        for x, y in zip(xs, ys):
            assert L.ref_bred_add(x, q, O._p(brc)) == x % q
            lz = L.ref_bred_add_lazy(x, q, O._p(brc))
            assert lz % q == x % q and lz < 2 * q
            assert L.ref_mform(x, q, O._p(brc)) == (x << 64) % q
            assert L.ref_imform(x, q, qinv) == x * rinv % q
            assert L.ref_bred(x, y, q, O._p(brc)) == x * y % q
            lz = L.ref_bred_lazy(x, y, q, O._p(brc))
            assert lz % q == x * y % q and lz < 2 * q
            yq = y % q                                  # MRed needs x*y < q*2^64
            assert L.ref_mred(x, yq, q, qinv) == x * yq * rinv % q
            lz = L.ref_mred_lazy(x, yq, q, qinv)
            assert lz % q == x * yq * rinv % q and lz < 2 * q
