for ln in 8 9 10 11 12; do python tools/profile_ntt.py --ring pow2 --logn $ln --polys $((1<<(22-ln))) --limbs 8 --iters 1 --time 2>&1 | grep -v "^    "; done
for a in 3 5 7 8 9 10 11 12; do python tools/profile_ntt.py --ring 3n --a $a --polys $((1<<(21-a))) --limbs 8 --iters 1 --time 2>&1 | grep -v "^    "; done
