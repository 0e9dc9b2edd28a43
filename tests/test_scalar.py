"""ring/scalar.go and the word-level primitives of ring/modular_reduction.go as restated for the host side (scalar.py):
against the oracle's C restatement (itself pinned by the reference's vectors) and against big integers."""
import numpy as np

from conftest import QI60
from b200ring import scalar as S
from oracle import oracle as O

MODULI = list(QI60[:3]) + [0x1000000000248001, 0x10000000000a1001, 4294967377]


def test_word_primitives_match_the_oracle_and_big_integers():
    L = O.lib()
    rng = np.random.default_rng(5)
    for q in MODULI:
        brc = O.bred_constant(q)
        assert S.BRedConstant(q) == (int(brc[0]), int(brc[1])) and S.MRedConstant(q) == L.ref_gen_mred_constant(q)
        qinv = S.MRedConstant(q)
        for x, y in zip(rng.integers(0, 1 << 64, 100, dtype=np.uint64), rng.integers(0, q, 100, dtype=np.uint64)):
            x, y = int(x), int(y)
            assert S.MRedLazy(x, y, q, qinv) == L.ref_mred_lazy(x, y, q, qinv)
            assert S.MRed(x, y, q, qinv) == L.ref_mred(x, y, q, qinv) == x * y * pow(1 << 64, -1, q) % q
            assert S.MForm(x, q) == L.ref_mform(x, q, O._p(brc)) == (x << 64) % q
            assert S.IMForm(x, q, qinv) == L.ref_imform(x, q, qinv)
            assert S.BRedAdd(x, q) == L.ref_bred_add(x, q, O._p(brc)) == x % q
        assert S.ModExp(3, q - 1, q) == 1


def test_rns_scalars():
    Q = MODULI
    big = 3 ** 200 + 12345
    a, b = S.NewRNSScalarFromBigint(Q, big), S.NewRNSScalarFromUInt64(Q, 0xfedcba9876543210)
    assert a == [big % q for q in Q] and b == [0xfedcba9876543210 % q for q in Q] and S.NewRNSScalar(Q) == [0] * len(Q)
    assert S.NewRNSScalarFromBigint(Q, -7) == [q - 7 for q in Q]
    assert S.NegRNSScalar(Q, a) == [q - x for x, q in zip(a, Q)] and S.NegRNSScalar(Q, [0] * len(Q)) == Q
    assert S.SubRNSScalar(Q, a, b) == [(x - y) % q for x, y, q in zip(a, b, Q)]
    am, bm = S.MFormRNSScalar(Q, a), S.MFormRNSScalar(Q, b)
    assert am == [x * (1 << 64) % q for x, q in zip(a, Q)]
    prod = S.MulRNSScalar(Q, am, bm)          # Montgomery form of a*b, lazy
    assert all(p < 2 * q and p % q == x * y * (1 << 64) % q for p, x, y, q in zip(prod, a, b, Q))
    inv = S.Inverse(Q, am)                    # Montgomery form of a^-1
    assert inv == [pow(x, -1, q) * (1 << 64) % q for x, q in zip(a, Q)]
    assert [S.MRed(i, m, q, S.MRedConstant(q)) for i, m, q in zip(inv, am, Q)] == [(1 << 64) % q for q in Q]
