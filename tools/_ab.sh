# A/B timing of tuning variants: usage tools/_ab.sh tag1 tag2 ... (empty tag "" = default library)
for t in "$@"; do
  lib=matrix-fhe-lattigo_b200/lib/libb200ring${t:+_$t}.so
  for r in 1 2; do
    echo "== ${t:-default} 3n"; B200RING_LIB=$PWD/$lib python tools/profile_ntt.py --ring 3n --polys 128 --limbs 24 --iters 1 --time 2>&1 | tail -2
  done
  echo "== ${t:-default} pow2"; B200RING_LIB=$PWD/$lib python tools/profile_ntt.py --ring pow2 --polys 128 --limbs 24 --iters 1 --time 2>&1 | tail -2
done
