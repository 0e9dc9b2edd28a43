#!/bin/bash
P="python tools/profile_ntt.py --polys 128 --time"
L=matrix-fhe-lattigo_b200/lib
{
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "3n or alternative or fold" 2>&1 | tail -2
echo "== default (split64)"; $P 2>&1 | tail -2
for v in nosplit rm3 rm5; do echo "== $v"; B200RING_LIB=$L/libb200ring_$v.so $P 2>&1 | tail -2; done
echo "== default again"; $P 2>&1 | tail -2
} > gpurun_out/r2q.txt 2>&1
cat gpurun_out/r2q.txt
