#!/bin/bash
# round 2 experiment A: pipe costs on real clocks, twiddle-load cost of the fold kernels, ncu of the full-size fold kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt
./tools/pipebench2 > gpurun_out/r2a_pipebench2.txt 2>&1
python tools/profile_ntt.py --polys 128 --time > gpurun_out/r2a_time_default.txt 2>&1
B200_SKIP_CHECK=1 B200RING_LIB=matrix-fhe-lattigo_b200/lib/libb200ring_twu.so python tools/profile_ntt.py --polys 128 --time > gpurun_out/r2a_time_twu.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_rows|k_cols" --launch-skip 8 -c 4 -f -o gpurun_out/r2a_fold python tools/profile_ntt.py --polys 128 --iters 4 > gpurun_out/r2a_ncu.log 2>&1
tail -3 gpurun_out/r2a_time_default.txt gpurun_out/r2a_time_twu.txt
head -30 gpurun_out/r2a_pipebench2.txt
