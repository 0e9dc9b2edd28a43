#!/bin/bash
ncu --set full --clock-control none --import-source on -k regex:"k_rows|k_cols" --launch-skip 8 -c 4 -f -o gpurun_out/r2m_fold python tools/profile_ntt.py --polys 128 --iters 4 > gpurun_out/r2m_ncu.log 2>&1
tail -2 gpurun_out/r2m_ncu.log
