# walk's best weight in a register; register budgets for multi-word rows now that the arena no longer bounds the occupancy
cp stabilizer_gs_b200/libsgs.so /tmp/libsgs_saved.so
for lib in /tmp/libsgs_saved.so build_ab/libsgs_minb5.so build_ab/libsgs_minb7.so; do
  cp $lib stabilizer_gs_b200/libsgs.so; echo "== $lib"
  for a in "24 300 0 4" "40 500 0 3"; do timeout 300 python scripts/run_sparse_once.py $a | tail -1 | sed 's/.*slow=/slow=/'; done
  timeout 300 python scripts/run_shard_once.py 40 1000 0 8 2 | head -1
done
cp /tmp/libsgs_saved.so stabilizer_gs_b200/libsgs.so
