#!/bin/bash
# round 2 experiment B: sensitivity of the fold kernels to occupancy, CTA size and the slot permutation
P="python tools/profile_ntt.py --polys 128 --time"
L=matrix-fhe-lattigo_b200/lib
run() { echo "== $1"; shift; env "$@" $P 2>&1 | tail -2; }
{
run default A=1
run tree A=1 TREEFLAG=1
run rows_3cta B200_ROW_SMEM_EXTRA_KB=40
run rows_2cta B200_ROW_SMEM_EXTRA_KB=80
run cols_6cta B200_COL_SMEM_EXTRA_KB=20
run cols_4cta B200_COL_SMEM_EXTRA_KB=40
run vt1 B200RING_LIB=$L/libb200ring_vt1.so
run minb5 B200RING_LIB=$L/libb200ring_minb5.so
run minb3 B200RING_LIB=$L/libb200ring_minb3.so
echo "== tree"; $P --tree 2>&1 | tail -2
} > gpurun_out/r2b.txt 2>&1
cat gpurun_out/r2b.txt
