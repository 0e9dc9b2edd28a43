"""Wire format of ring.Poly and of the ring parameters literal.

Reference: ring/poly.go:97-160 (Poly = structs.Matrix[uint64]), utils/structs/matrix.go:57-153,
utils/structs/vector.go:43-102, ring/ring.go:560-657, ring/subring.go:330-417.  The reference's own tests only
round-trip (ring/ring_test.go:333-430), so the byte layout is pinned here by vectors written out by hand from those
definitions; the CUDA-side marshalling is then compared with the oracle on the GPU.
"""
import io
import json

import numpy as np
import pytest

from oracle import oracle as O

KAT_Q = [576460752303439873, 576460752303702017]   # ring/ntt_test.go:10-89

# 2 limbs x 4 words, written out by hand: rows=2 | len=4 | 0,1,2,3 | len=4 | 2^64-1, 2^32, 0x0102030405060708, 7
HAND = bytes.fromhex(
    "0200000000000000"
    "0400000000000000" "0000000000000000" "0100000000000000" "0200000000000000" "0300000000000000"
    "0400000000000000" "ffffffffffffffff" "0000000001000000" "0807060504030201" "0700000000000000")
HAND_ROWS = [[0, 1, 2, 3], [2**64 - 1, 2**32, 0x0102030405060708, 7]]


def test_oracle_poly_bytes_match_hand_vector():
    assert O.poly_marshal([np.array(r, dtype=np.uint64) for r in HAND_ROWS]) == HAND
    rows, used = O.poly_unmarshal(HAND)
    assert used == len(HAND) == 8 + 2 * (8 + 4 * 8)       # Matrix.BinarySize: 8 + sum(8 + 8*len)
    assert [list(map(int, r)) for r in rows] == HAND_ROWS


def test_oracle_poly_bytes_ragged_and_empty():
    assert O.poly_marshal([]) == (0).to_bytes(8, "little")
    d = O.poly_marshal([np.array([], dtype=np.uint64), np.array([5], dtype=np.uint64)])
    assert d == bytes.fromhex("0200000000000000" "0000000000000000" "0100000000000000" "0500000000000000")
    rows, used = O.poly_unmarshal(d)
    assert used == len(d) and [len(r) for r in rows] == [0, 1]
    with pytest.raises(ValueError):
        O.poly_unmarshal(d[:-1])


def test_oracle_ring_literal():
    """Go's encoding/json of ringParametersLiteral: declaration-order keys, no spaces."""
    data = O.ring_literal_json(16, KAT_Q, 2)
    lit = json.loads(data)
    assert data.startswith(b'[{"Type":0,"LogN":4,"NthRoot":2,"Modulus":576460752303439873,"Factors":[2,3,')
    assert b" " not in data
    for p, q in zip(lit, KAT_Q):
        assert list(p) == ["Type", "LogN", "NthRoot", "Modulus", "Factors", "PrimitiveRoot"]
        m = q - 1
        for f in p["Factors"]:   # unique prime factors of q-1, ascending, complete
            assert m % f == 0 and all(f % d for d in range(2, min(f, 1000)))
            while m % f == 0:
                m //= f
        assert m == 1 and p["Factors"] == sorted(p["Factors"])
        g = p["PrimitiveRoot"]   # smallest primitive root >= 3 (subring.go:236-248)
        isroot = lambda x: all(pow(x, (q - 1) // f, q) != 1 for f in p["Factors"])
        assert isroot(g) and not any(isroot(x) for x in range(3, g))
    # 3N literal: LogN = bits.Len64(N-1), NthRoot = 3, PrimitiveRoot = omega of order 3N
    q3 = O.find_3n_rns_primes(12, 60, 1)[0]
    p = json.loads(O.ring_literal_json(12, [q3], 3))[0]
    assert (p["Type"], p["LogN"], p["NthRoot"]) == (0, 4, 3)
    w = p["PrimitiveRoot"]
    assert pow(w, 36, q3) == 1 and pow(w, 18, q3) != 1 and pow(w, 12, q3) != 1


# ---- product path (GPU) --------------------------------------------------------------------------------
def _ring(N, moduli):
    import b200ring
    return b200ring, b200ring.NewRing(N, moduli)


@pytest.mark.gpu
def test_poly_marshal_matches_oracle_and_hand_vector():
    b200ring, r = _ring(16, KAT_Q)
    rng = np.random.default_rng(5)
    data = np.stack([rng.integers(0, q, 16, dtype=np.uint64) for q in KAT_Q])
    p = b200ring.NewPoly(r).set(data)
    want = O.poly_marshal(list(data))
    assert p.BinarySize() == len(want) == 8 + 2 * (8 + 16 * 8)
    assert p.MarshalBinary() == want
    assert p.MarshalBinary(level=0) == O.poly_marshal([data[0]])          # AtLevel(0) view of the same poly
    buf = io.BytesIO()
    assert p.WriteTo(buf) == len(want) and buf.getvalue() == want       # ring_test.go:402-425
    p2 = b200ring.NewPoly(r)
    buf.seek(0)
    assert p2.ReadFrom(buf) == len(want) and p2.Equal(p)
    # hand vector on a 4-word... the ring degree is 16, so embed the hand rows check through a batch-2 poly instead
    pb = b200ring.NewPoly(r, batch=2)
    assert pb.UnmarshalBinary(want, b=1) == (1, len(want))
    got = pb.get()
    assert (got[1] == data).all() and (got[0] == 0).all()
    assert pb.MarshalBinary(b=1) == want and pb.MarshalBinary(b=0) == O.poly_marshal([np.zeros(16, np.uint64)] * 2)


@pytest.mark.gpu
def test_poly_unmarshal_partial_level_and_errors():
    b200ring, r = _ring(16, KAT_Q)
    base = np.stack([np.arange(16, dtype=np.uint64) + 100 * i for i in range(2)])
    p = b200ring.NewPoly(r).set(base)
    one = O.poly_marshal([np.full(16, 9, dtype=np.uint64)])
    assert p.UnmarshalBinary(one + b"trailing") == (0, len(one))          # consumes exactly one polynomial
    got = p.get()[0]
    assert (got[0] == 9).all() and (got[1] == base[1]).all()              # limbs above the level are untouched
    before = p.get().copy()
    for bad in (one[:7], one[:20], one[:-1],                              # truncated count / header / data
                O.poly_marshal([np.zeros(8, np.uint64)]),                 # wrong limb length
                O.poly_marshal([np.zeros(16, np.uint64)] * 3)):           # more limbs than moduli
        with pytest.raises(b200ring.B200Error) as e:
            p.UnmarshalBinary(bad)
        assert e.value.code == -4
    assert (p.get() == before).all()                                      # failed reads leave the polynomial unchanged
    assert p.UnmarshalBinary(O.poly_marshal([])) == (-1, 8)               # an empty Matrix is legal


@pytest.mark.gpu
def test_ring_literal_and_roundtrip():
    import b200ring
    from conftest import QI60
    r = b200ring.NewRing(1 << 10, QI60[:3])
    data = r.MarshalJSON()
    assert data == O.ring_literal_json(1 << 10, QI60[:3], 2) == r.MarshalBinary()
    r2 = b200ring.UnmarshalJSON(data)                                     # ring_test.go:335-350
    assert (r2.N(), r2.ModuliChain(), r2.Type()) == (r.N(), r.ModuliChain(), r.Type())
    for a, b in zip(r.SubRings, r2.SubRings):
        assert a.PrimitiveRoot == b.PrimitiveRoot and (a.RootsForward == b.RootsForward).all()
    # 3N ring: same bytes as the reference's literal; N is not recoverable from it (LogN only)
    N = 3 << 6
    q3 = O.find_3n_rns_primes(N, 60, 2)
    m = b200ring.NewRing(N, q3)
    d3 = m.MarshalJSON()
    assert d3 == O.ring_literal_json(N, q3, 3)
    with pytest.raises(ValueError):
        b200ring.UnmarshalJSON(d3)
    m2 = b200ring.UnmarshalJSON(d3, N=N)
    assert m2.Type() == b200ring.Matrix and [s.Root for s in m2.SubRings] == [s.Root for s in m.SubRings]
    with pytest.raises(ValueError):
        b200ring.UnmarshalJSON(d3, N=2 * N)
    lit = json.loads(data)
    lit[1]["LogN"] = 11
    with pytest.raises(ValueError, match="same ring degree"):
        b200ring.UnmarshalJSON(json.dumps(lit).encode())
    lit = json.loads(data)
    lit[0]["Type"] = 3
    with pytest.raises(ValueError, match="invalid ring type"):
        b200ring.UnmarshalJSON(json.dumps(lit).encode())
