(time timeout 900 python -m pytest tests -m gpu -x -q -k "not config5") > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2e_tests.log
for i in 1 2; do python scripts/run_sparse_once.py 24 300 0 5 | tail -1; python scripts/run_sparse_once.py 40 500 0 3 | tail -1; done
timeout 100 python scripts/run_sparse_once.py 40 1000 0 1 | tail -1
echo "== default"; python scripts/shard_sweep.py 24 300 0 8 3 | tail -1
echo "== NOSPLIT"; SGS_NOSPLIT=1 python scripts/shard_sweep.py 24 300 0 8 3 | tail -1
echo "== world 2/4"; python scripts/shard_sweep.py 24 300 0 2 3 | tail -1; python scripts/shard_sweep.py 24 300 0 4 3 | tail -1
cp build_ab/libsgs_trace.so stabilizer_gs_b200/libsgs.so
echo "== trace build, world 8"; SGS_TRACE=1 python scripts/shard_sweep.py 24 300 0 8 2 2>&1 | grep -A2 "replay shard 3" | tail -3
SGS_TRACE=1 SGS_NOSPLIT=1 python scripts/shard_sweep.py 24 300 0 8 2 2>&1 | grep -A2 "replay shard 3" | tail -3
