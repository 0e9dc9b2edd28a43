#!/bin/bash
P="python tools/profile_ntt.py --polys 128 --time"
{
for cw in 4 8 16; do echo "== bal CW=$cw"; B200_CW_BAL=$cw $P 2>&1 | tail -2; done
} > gpurun_out/r2j.txt 2>&1
cat gpurun_out/r2j.txt
