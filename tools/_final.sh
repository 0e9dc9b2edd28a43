set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r1c.json 2> gpurun_out/bench_ref_r1c.err; tail -c 300 gpurun_out/bench_ref_r1c.json
python bench.py > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; cat gpurun_out/bench_r1c.json
B200_PROFILE_NEXT=1 python tools/profile_ntt.py --ring 3n --polys 64 --limbs 24 --iters 1 2>&1 | tail -6
