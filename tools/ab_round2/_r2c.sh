#!/bin/bash
P="python tools/profile_ntt.py --polys 128 --time"
L=matrix-fhe-lattigo_b200/lib
{
for v in inv4 fwd5 cw8; do echo "== $v"; B200RING_LIB=$L/libb200ring_$v.so $P 2>&1 | tail -2; done
} > gpurun_out/r2c.txt 2>&1
cat gpurun_out/r2c.txt
