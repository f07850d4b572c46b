#!/bin/bash
# A/B two builds of libsgs.so on the same box: scripts/ab_lib.sh libA.so libB.so [n T]   (default 24 300)
cd "$(dirname "$0")/.."
cp stabilizer_gs_b200/libsgs.so /tmp/libsgs_saved.so
LIBS=(); ARGS=()
for a in "$@"; do case "$a" in *.so) LIBS+=("$a");; *) ARGS+=("$a");; esac; done
N=${ARGS[0]:-24}; T=${ARGS[1]:-300}
for round in 1 2; do
  for lib in "${LIBS[@]}"; do
    cp "$lib" stabilizer_gs_b200/libsgs.so
    echo "$lib: $(python scripts/run_sparse_once.py $N $T 0 4 | tail -1 | sed 's/.*dev=/dev=/')"
  done
done
cp /tmp/libsgs_saved.so stabilizer_gs_b200/libsgs.so
