#!/bin/bash
P="python tools/profile_ntt.py --polys 128 --time"
{
echo "== default"; $P 2>&1 | tail -2
for r in 50 60 75 100; do echo "== rows carveout $r"; B200_CARVEOUT_ROWS=$r $P 2>&1 | tail -1; done
for c in 70 80 100; do echo "== cols carveout $c"; B200_CARVEOUT_COLS=$c $P 2>&1 | tail -1; done
} > gpurun_out/r2v.txt 2>&1
cat gpurun_out/r2v.txt
