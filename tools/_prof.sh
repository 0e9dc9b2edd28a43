ncu --set full --clock-control none --import-source on -k regex:"inv" -c 2 -f -o gpurun_out/prof_final_inv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-verify > gpurun_out/ncu_f2.log 2>&1
ls -la gpurun_out/prof_final_inv.ncu-rep
