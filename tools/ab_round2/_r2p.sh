#!/bin/bash
P="python tools/profile_ntt.py --polys 128 --time"
{
python -m pytest tests/test_gpu_parity.py tests/test_multigpu.py tests/test_keyswitch.py -x -q -m gpu 2>&1 | tail -3
echo "== pow2 2^16 (Shoup, QI60)"; $P --ring pow2 2>&1 | tail -2
echo "== 3n shoup (NO_FOLD)"; B200_NO_FOLD=1 $P 2>&1 | tail -2
echo "== 3n default"; $P 2>&1 | tail -2
} > gpurun_out/r2p.txt 2>&1
cat gpurun_out/r2p.txt
