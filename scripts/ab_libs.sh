#!/bin/bash
# A/B several builds of libsgs.so on the same box over the sparse workloads: scripts/ab_libs.sh lib1.so lib2.so ...
cd "$(dirname "$0")/.."
cp stabilizer_gs_b200/libsgs.so /tmp/libsgs_saved.so
for lib in /tmp/libsgs_saved.so "$@"; do
  cp $lib stabilizer_gs_b200/libsgs.so; echo "== $lib"
  for a in "24 300 0 5" "30 400 0 3" "40 500 0 3"; do timeout 300 python scripts/run_sparse_once.py $a | tail -1 | sed 's/.*slow=/slow=/'; done
  timeout 300 python scripts/run_shard_once.py 40 1000 0 8 2 2>/dev/null | head -1
done
cp /tmp/libsgs_saved.so stabilizer_gs_b200/libsgs.so
