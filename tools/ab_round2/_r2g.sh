#!/bin/bash
P="python tools/profile_ntt.py --polys 128 --time"
L=matrix-fhe-lattigo_b200/lib
{
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
echo "== default (bal MINB17)"; $P 2>&1 | tail -2
echo "== bal off"; B200_COLS_BAL=0 $P 2>&1 | tail -2
echo "== pow2"; $P --ring pow2 2>&1 | tail -2
} > gpurun_out/r2g.txt 2>&1
cat gpurun_out/r2g.txt
