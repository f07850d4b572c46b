# warp end-time histogram of the single-GPU walk, with and without the heavy-first order
cp stabilizer_gs_b200/libsgs.so /tmp/libsgs_saved.so
cp build_ab/libsgs_trace.so stabilizer_gs_b200/libsgs.so
echo "== trace build, config 4"; SGS_TRACE=1 timeout 120 python scripts/run_sparse_once.py 24 300 0 3 2>&1 | grep -v "level\|work stack" | tail -7
echo "== trace build, config 4, SGS_PERM=1"; SGS_PERM=1 SGS_TRACE=1 timeout 120 python scripts/run_sparse_once.py 24 300 0 3 2>&1 | grep -v "level\|work stack" | tail -7
echo "== trace build, n=30 T=400"; SGS_TRACE=1 timeout 120 python scripts/run_sparse_once.py 30 400 0 2 2>&1 | grep -v "level\|work stack" | tail -7
cp /tmp/libsgs_saved.so stabilizer_gs_b200/libsgs.so
echo "== normal build"; timeout 120 python scripts/run_sparse_once.py 24 300 0 4 | tail -2
echo "== normal build SGS_PERM=1"; SGS_PERM=1 timeout 120 python scripts/run_sparse_once.py 24 300 0 4 | tail -2
