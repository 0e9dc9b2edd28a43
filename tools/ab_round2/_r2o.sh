#!/bin/bash
{
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "polymul or product or mul" 2>&1 | tail -3
B200_PROFILE_MUL=1 python tools/profile_ntt.py --polys 64 --limbs 16 2>&1 | tail -2
B200_FOLD_MUL=0 B200_PROFILE_MUL=1 python tools/profile_ntt.py --polys 64 --limbs 16 2>&1 | tail -2
} > gpurun_out/r2o.txt 2>&1
cat gpurun_out/r2o.txt
