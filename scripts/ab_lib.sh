#!/bin/bash
# A/B two builds of libsgs.so on the same box: scripts/ab_lib.sh libA.so libB.so
cd "$(dirname "$0")/.."
cp stabilizer_gs_b200/libsgs.so /tmp/libsgs_saved.so
for round in 1 2; do
  for lib in "$@"; do
    cp "$lib" stabilizer_gs_b200/libsgs.so
    echo "$lib: $(python scripts/run_sparse_once.py 24 300 0 5 | tail -1 | sed 's/.*dev=/dev=/')"
  done
done
cp /tmp/libsgs_saved.so stabilizer_gs_b200/libsgs.so
