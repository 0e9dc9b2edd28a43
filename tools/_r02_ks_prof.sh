#!/bin/bash
# round 2: launch list and ncu --set full summaries of the key switch and of the fused ring product (one GPU)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_keyswitch.csv python tools/bench_keyswitch.py --reps 2 > gpurun_out/r02_ks_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_basis_extend|k_gadget_mac" --launch-skip 6 -c 2 -f -o gpurun_out/r02_ks_kernels python tools/bench_keyswitch.py --reps 2 > gpurun_out/r02_ks_f.log 2>&1
B200_PROFILE_MUL=1 ncu --set full --clock-control none --import-source on -k regex:"k_rows_mul" --launch-skip 2 -c 1 -f -o gpurun_out/r02_mul_kernel python tools/profile_ntt.py --polys 64 --limbs 16 --iters 1 > gpurun_out/r02_mul_f.log 2>&1
tail -2 gpurun_out/r02_ks_f.log gpurun_out/r02_mul_f.log
