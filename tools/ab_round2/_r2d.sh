#!/bin/bash
P="python tools/profile_ntt.py --polys 128 --time"
L=matrix-fhe-lattigo_b200/lib
{
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
echo "== bal default"; $P 2>&1 | tail -2
echo "== bal off"; B200_COLS_BAL=0 $P 2>&1 | tail -2
for v in balu2 bal14 bal12; do echo "== $v"; B200RING_LIB=$L/libb200ring_$v.so $P 2>&1 | tail -2; done
} > gpurun_out/r2d.txt 2>&1
cat gpurun_out/r2d.txt
