#!/bin/bash
P="python tools/profile_ntt.py --polys 64 --limbs 8 --time"
{
python -m pytest tests/test_gpu_parity.py tests/test_keyswitch.py tests/test_multigpu.py -x -q -m gpu 2>&1 | tail -3
for a in 12 13 14 16; do echo "== 3n a=$a"; $P --a $a 2>&1 | tail -2; echo "-- nofoldcols"; B200_FOLD_COLS=0 $P --a $a 2>&1 | tail -2 | head -1; done
echo "== default headline"; python tools/profile_ntt.py --polys 128 --time 2>&1 | tail -2
} > gpurun_out/r2r.txt 2>&1
cat gpurun_out/r2r.txt
