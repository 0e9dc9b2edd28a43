#!/usr/bin/env python3
"""Generate golden vectors for the fast 3N transform from the REFERENCE's own Python prototype
(references/integer_dft.py: IntegerDFT.factorized_dft / factorized_idft) -> ntt3n_proto.json.

Run in the build container only (imports from /root/reference, absent on the GPU box):
    python tests/golden/make_ntt3n_proto.py
The prototype enumerates all of Z_p at construction, so the primes are small (~2^16..2^20); the
vectors pin the layer structure, twiddle tree and slot order, not the 61-bit arithmetic.
"""
import json
import os
import random
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/references"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ntt3n_proto.json")
sys.path.insert(0, REF)
from integer_dft import IntegerDFT  # noqa: E402


def main():
    rng = random.Random(0xB200)
    vectors = []
    for N, bits in [(6, 10), (12, 12), (18, 12), (24, 14), (36, 14), (48, 16), (54, 16), (96, 16), (192, 17)]:
        d = IntegerDFT(N, min_bits=bits)
        p, w = int(d.p), int(d.w)
        coeffs = [rng.randrange(p) for _ in range(N)]
        out = [int(x) for x in d.factorized_dft(coeffs)]
        back = [int(x) for x in d.factorized_idft(out)]
        assert back == coeffs
        leaf = [int(x) % (3 * N) for x in d.tree[d.level]]
        # sanity: slot i holds f(w^leaf[i])
        for i in (0, N // 2, N - 1):
            x = pow(w, leaf[i], p)
            assert sum(c * pow(x, j, p) for j, c in enumerate(coeffs)) % p == out[i]
        vectors.append({"N": N, "p": p, "w": w, "zetas": [int(z) for z in d.zetas],
                        "leaf_exponents": leaf, "coeffs": coeffs, "ntt_tree_order": out})
        print("N=%d p=%d w=%d" % (N, p, w))
    with open(OUT, "w") as f:
        json.dump({"source": "references/integer_dft.py IntegerDFT.factorized_dft/idft (tree order)",
                   "vectors": vectors}, f)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
