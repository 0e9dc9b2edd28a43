#!/bin/bash
# round 2: the numbers and captures profiles/ quotes (one GPU)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02_pytest_gpu.txt
python __graft_entry__.py --smoke 2>&1 | tail -2 >> gpurun_out/r02_pytest_gpu.txt
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
python tools/bench_configs.py --out gpurun_out/r02_configs.json > gpurun_out/r02_configs.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_rows|k_cols" --launch-skip 40 -c 4 -f -o gpurun_out/r02_bench_kernels python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-verify > gpurun_out/r02_ncu_f.log 2>&1
./tools/solbench > gpurun_out/r02_solbench.txt 2>&1
./tools/pipebench2 > gpurun_out/r02_pipebench2.txt 2>&1
cat gpurun_out/r02_pytest_gpu.txt; cat gpurun_out/r02_bench.json | cut -c1-600; tail -3 gpurun_out/r02_ncu_f.log
ncu --set full --clock-control none --import-source on -k regex:"_inv_" --launch-skip 10 -c 2 -f -o gpurun_out/r02_bench_kernels_inv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-verify > gpurun_out/r02_ncu_f2.log 2>&1
