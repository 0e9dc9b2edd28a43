#!/usr/bin/env python3
"""Extract the known-answer vectors of the reference's ring/ntt_test.go:10-89 into ntt_kat.json.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_ntt_kat.py
The JSON it writes is the committed fixture; tests never read /root/reference.
"""
import json
import os
import re
import sys

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/ring/ntt_test.go"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ntt_kat.json")


def main():
    text = open(SRC).read()
    body = text[text.index("var testVector"): text.index("func TestNTT")]
    # each entry: N, []uint64{q...}, Poly{[][]uint64{ {..},{..} }}, Poly{[][]uint64{ {..},{..} }}
    entries = []
    pos = 0
    pat = re.compile(r"\{\s*(\d+),\s*\[\]uint64\{([^}]*)\},", re.S)
    while True:
        m = pat.search(body, pos)
        if not m:
            break
        n = int(m.group(1))
        qis = [int(x) for x in m.group(2).replace("\n", " ").split(",") if x.strip()]
        pos = m.end()
        polys = []
        for _ in range(2):
            start = body.index("Poly{[][]uint64{", pos) + len("Poly{[][]uint64{")
            end = body.index("}}", start)
            limbs = re.findall(r"\{([^{}]*)\}", body[start:end + 1])
            polys.append([[int(x) for x in l.replace("\n", " ").split(",") if x.strip()] for l in limbs])
            pos = end + 2
        assert all(len(l) == n for p in polys for l in p), n
        assert len(polys[0]) == len(qis) == len(polys[1])
        entries.append({"N": n, "Qis": qis, "poly": polys[0], "polyNTT": polys[1]})
    assert [e["N"] for e in entries] == [16, 32, 64, 128, 256, 512], [e["N"] for e in entries]
    with open(OUT, "w") as f:
        json.dump({"source": "ring/ntt_test.go:10-89 (testVector)", "vectors": entries}, f)
    print("wrote", OUT, [e["N"] for e in entries])


if __name__ == "__main__":
    main()
