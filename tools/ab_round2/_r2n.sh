#!/bin/bash
P="python tools/profile_ntt.py --polys 128 --time"
{
for mb in 256 1024 2304 4096; do echo "== CHUNK_MB=$mb"; B200_CHUNK_MB=$mb $P 2>&1 | tail -2; done
python tools/bench_configs.py --out gpurun_out/r02_configs.json 2>&1 | tail -8
python tools/bench_keyswitch.py 2>&1 | tail -4
} > gpurun_out/r2n.txt 2>&1
cat gpurun_out/r2n.txt
