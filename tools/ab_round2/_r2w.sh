#!/bin/bash
P="python tools/profile_ntt.py --polys 128 --time"
L=matrix-fhe-lattigo_b200/lib
{
echo "== stream data"; B200RING_LIB=$L/libb200ring_sd1.so $P 2>&1 | tail -2
echo "== default"; $P 2>&1 | tail -2
} > gpurun_out/r2w.txt 2>&1
cat gpurun_out/r2w.txt
