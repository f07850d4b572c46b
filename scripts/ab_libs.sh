#!/bin/bash
# same-box A/B of several builds: scripts/ab_libs.sh build_ab/libsgs_a.so build_ab/libsgs_b.so ...  (2 rounds each)
cd "$(dirname "$0")/.."
cp stabilizer_gs_b200/libsgs.so /tmp/libsgs_saved.so
for round in 1 2; do
  for lib in "$@"; do
    cp "$lib" stabilizer_gs_b200/libsgs.so
    echo "$lib c4 : $(python scripts/run_sparse_once.py 24 300 0 5 | tail -1 | sed 's/.*cand=/cand=/')"
    echo "$lib w2 : $(python scripts/run_sparse_once.py 40 500 0 3 | tail -1 | sed 's/.*cand=/cand=/')"
    echo "$lib w2b: $(python scripts/run_sparse_once.py 40 700 0 2 | tail -1 | sed 's/.*cand=/cand=/')"
  done
done
cp /tmp/libsgs_saved.so stabilizer_gs_b200/libsgs.so
