#!/bin/bash
P="python tools/profile_ntt.py --polys 128 --time"
L=matrix-fhe-lattigo_b200/lib
{
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
echo "== separate staging"; $P 2>&1 | tail -2
echo "== shared staging (stg0)"; B200RING_LIB=$L/libb200ring_stg0.so $P 2>&1 | tail -2
echo "== separate again"; $P 2>&1 | tail -2
} > gpurun_out/r2u.txt 2>&1
cat gpurun_out/r2u.txt
