(time timeout 900 python -m pytest tests -m gpu -x -q -k "not config5") > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2l_tests.log
for i in 1 2; do python scripts/run_sparse_once.py 24 300 0 5 | tail -1; SGS_NO_DFS_EXPAND=1 python scripts/run_sparse_once.py 24 300 0 5 | tail -1; done
python scripts/run_sparse_once.py 40 500 0 3 | tail -1
timeout 60 python scripts/run_sparse_once.py 40 1000 0 1 | tail -1
python scripts/shard_sweep.py 24 300 0 8 3 | tail -1
SGS_TRACE=1 python scripts/run_shard_once.py 24 300 0 8 3 2>&1 | grep "replay shard" | tail -1
python scripts/e2e_trace.py 2>&1 | grep one-shot | tail -2
