#!/bin/bash
P="python tools/profile_ntt.py --polys 128 --time"
{
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
echo "== default"; $P 2>&1 | tail -2
echo "== pow2 N=2^15 (fold rows? 61-bit primes: shoup)"; $P --ring pow2 --logn 15 2>&1 | tail -2
echo "== 3n a=14"; $P --a 14 2>&1 | tail -2
} > gpurun_out/r2k.txt 2>&1
cat gpurun_out/r2k.txt
