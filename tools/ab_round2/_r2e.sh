#!/bin/bash
P="python tools/profile_ntt.py --polys 128 --time"
L=matrix-fhe-lattigo_b200/lib
{
echo "== bal default (unroll 1)"; $P 2>&1 | tail -2
echo "== bal14"; B200RING_LIB=$L/libb200ring_bal14.so $P 2>&1 | tail -2
echo "== ws, bal off"; B200_COLS_BAL=0 B200RING_LIB=$L/libb200ring_ws.so $P 2>&1 | tail -2
echo "== ws, bal on"; B200RING_LIB=$L/libb200ring_ws.so $P 2>&1 | tail -2
} > gpurun_out/r2e.txt 2>&1
cat gpurun_out/r2e.txt
