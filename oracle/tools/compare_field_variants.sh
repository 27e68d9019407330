#!/bin/bash
# Times the C oracle's field-layer variants (dedicated squaring, lazily reduced Fp2 product) on THIS machine's CPU, so that the
# CPU baseline is built with the fastest one (a baseline must not be made slower by "fidelity").  TEST INFRASTRUCTURE ONLY.
cd "$(dirname "$0")/.."
T=${1:-16}
for defs in "" "-DORACLE_SQR_DEDICATED" "-DORACLE_FP2_LAZY" "-DORACLE_SQR_DEDICATED -DORACLE_FP2_LAZY"; do
  for flags in "-O3 -march=native" "-O3"; do
    mkdir -p _build
    gcc $flags $defs -fPIC -Wno-unused-function -shared -o _build/libbls_oracle_c.so bls_oracle_c.c 2>&1 | head -3
    (cd .. && python -c "
from oracle import cpu_baseline as c
r=c.run(32*$T,128,threads=$T); r=c.run(32*$T,128,threads=$T); print('[$defs]', '$flags', round(32*$T/r['seconds']), 'sets/s', end='')
h=c.run_hash_to_g2(256*$T,$T); print('   hash_to_G2', round(h['msgs']/h['seconds']), '/s')
")
  done
done
make -s clean; make -s
