#!/bin/bash
P="python tools/profile_ntt.py --polys 128 --time"
L=matrix-fhe-lattigo_b200/lib
{
B200_ROWS_HALF=1 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
echo "== half (MINB 5)"; B200_ROWS_HALF=1 $P 2>&1 | tail -2
for v in h4 h6; do echo "== half $v"; B200_ROWS_HALF=1 B200RING_LIB=$L/libb200ring_$v.so $P 2>&1 | tail -2; done
echo "== default (full warp rows)"; $P 2>&1 | tail -2
echo "== half tree"; B200_ROWS_HALF=1 $P --tree 2>&1 | tail -2
echo "== polymul half / full"; B200_ROWS_HALF=1 B200_PROFILE_MUL=1 python tools/profile_ntt.py --polys 64 --limbs 16 2>&1 | tail -2;  B200_PROFILE_MUL=1 python tools/profile_ntt.py --polys 64 --limbs 16 2>&1 | tail -2
} > gpurun_out/r2t.txt 2>&1
cat gpurun_out/r2t.txt
