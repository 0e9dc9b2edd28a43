#!/bin/bash
P="python tools/profile_ntt.py --polys 128 --time"
L=matrix-fhe-lattigo_b200/lib
{
for v in vt4 vt3; do echo "== $v"; B200RING_LIB=$L/libb200ring_$v.so $P 2>&1 | tail -2; done
echo "== default"; $P 2>&1 | tail -2
} > gpurun_out/r2s.txt 2>&1
cat gpurun_out/r2s.txt
