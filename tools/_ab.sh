python -m pytest tests/test_gpu_parity.py tests/test_keyswitch.py tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -2
python tools/profile_ntt.py --ring 3n --polys 128 --iters 1 --time
python tools/profile_ntt.py --ring pow2 --polys 128 --iters 1 --time
for ln in 10 12 14 17; do python tools/profile_ntt.py --ring pow2 --logn $ln --polys $((1<<(22-ln))) --limbs 8 --iters 1 --time 2>&1 | grep -v "^    "; done
for a in 10 12 14 16; do python tools/profile_ntt.py --ring 3n --a $a --polys $((1<<(21-a))) --limbs 8 --iters 1 --time 2>&1 | grep -v "^    "; done
B200_PROFILE_MUL=1 python tools/profile_ntt.py --ring 3n --polys 64 --limbs 16 --iters 1 2>&1 | tail -2
