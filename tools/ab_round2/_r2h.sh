#!/bin/bash
P="python tools/profile_ntt.py --polys 128 --time"
L=matrix-fhe-lattigo_b200/lib
{
echo "== default (bal MINB12)"; $P 2>&1 | tail -2
for v in bu2 bu3; do echo "== $v"; B200RING_LIB=$L/libb200ring_$v.so $P 2>&1 | tail -2; done
} > gpurun_out/r2h.txt 2>&1
cat gpurun_out/r2h.txt
