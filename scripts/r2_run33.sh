cp stabilizer_gs_b200/libsgs.so /tmp/libsgs_saved.so
for lib in /tmp/libsgs_saved.so build_ab/libsgs_h0r1.so build_ab/libsgs_h1r0.so build_ab/libsgs_h0r0.so; do
  cp $lib stabilizer_gs_b200/libsgs.so; echo "== $lib"
  for a in "24 300 0 5" "30 400 0 3"; do timeout 300 python scripts/run_sparse_once.py $a | tail -2 | sed 's/.*slow=/slow=/'; done
done
cp /tmp/libsgs_saved.so stabilizer_gs_b200/libsgs.so
