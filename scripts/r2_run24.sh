# where is the idle time of the single-GPU walk? trace build (per-root times, warp lifetimes) + forced hand-off rounds
cp stabilizer_gs_b200/libsgs.so /tmp/libsgs_saved.so
cp build_ab/libsgs_trace.so stabilizer_gs_b200/libsgs.so
echo "== trace build, config 4"; SGS_TRACE=1 timeout 120 python scripts/run_sparse_once.py 24 300 0 3 2>&1 | grep -v "level\|work stack" | tail -12
echo "== trace build, config 4, SGS_SPLIT=1"; SGS_SPLIT=1 SGS_TRACE=1 timeout 120 python scripts/run_sparse_once.py 24 300 0 3 2>&1 | grep -v "level\|work stack" | tail -6
echo "== trace build, n=40 T=500"; SGS_TRACE=1 timeout 120 python scripts/run_sparse_once.py 40 500 0 2 2>&1 | grep -v "level\|work stack" | tail -6
cp /tmp/libsgs_saved.so stabilizer_gs_b200/libsgs.so
echo "== normal build"; timeout 120 python scripts/run_sparse_once.py 24 300 0 4 | tail -2
echo "== normal build SGS_SPLIT=1"; SGS_SPLIT=1 timeout 120 python scripts/run_sparse_once.py 24 300 0 4 | tail -2
for sc in 5000 10000 20000; do echo "== SGS_SPLIT=1 SGS_SPLIT_CLIQUES=$sc"; SGS_SPLIT=1 SGS_SPLIT_CLIQUES=$sc timeout 120 python scripts/run_sparse_once.py 24 300 0 4 | tail -1; done
