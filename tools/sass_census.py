#!/usr/bin/env python3
"""Static SASS instruction census of the transform kernels (no GPU needed): `cuobjdump -sass` of the built object,
opcode counts per kernel.  The row kernels are one warp per 256-word row with 8 words per lane, i.e. 32 butterflies
per lane, so count / 32 is instructions per butterfly (both sides of uniform branches are counted).

    python tools/sass_census.py [object-or-library] [kernel-name-substring ...]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "matrix-fhe-lattigo_b200", "csrc", "_obj", "ntt_big.o")
want = sys.argv[2:] or ["k_rows_fwdILb1ELi2", "k_rows_invILb1ELi2", "k_cols_fwd_m3_15ILi4", "k_cols_inv_m3_15ILi4ELb0",
                        "k_rows_fwdILb0ELi4", "k_rows_invILb0ELi4"]
sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
cur, cnt = None, {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        cnt[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(2)
        cnt[cur][op.split(".")[0] + (".WIDE" if ".WIDE" in op else "")] += 1
for name, c in cnt.items():
    if any(w in name for w in want):
        short = re.sub(r"^_ZN\d+_GLOBAL__N__[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]+\d+", "", name)
        print("%s: %d instructions" % (short, sum(c.values())))
        print("    " + ", ".join("%s %d" % kv for kv in c.most_common(24)))
