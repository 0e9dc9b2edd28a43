for t in "$@"; do
  lib=matrix-fhe-lattigo_b200/lib/libb200ring${t:+_$t}.so
  echo "== ${t:-default}"; B200RING_LIB=$PWD/$lib python tools/bench_keyswitch.py --profile 2>&1 | tail -2 | cut -c1-330
done
