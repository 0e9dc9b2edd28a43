# one 3N timing per variant library: tools/_ab3.sh tag...
for t in "$@"; do
  lib=matrix-fhe-lattigo_b200/lib/libb200ring${t:+_$t}.so
  echo "== ${t:-default}"; B200RING_LIB=$PWD/$lib python tools/profile_ntt.py --ring 3n --polys 128 --limbs 24 --iters 1 --time 2>&1 | tail -2
done
