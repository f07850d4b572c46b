(time timeout 900 python -m pytest tests -m gpu -x -q -k "not config5") > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2j_tests.log
for i in 1 2; do python scripts/run_sparse_once.py 24 300 0 5 | tail -1; done
python scripts/shard_sweep.py 24 300 0 8 3 | tail -1
SGS_TRACE=1 python scripts/run_shard_once.py 24 300 0 8 3 2>&1 | grep "replay shard" | tail -2
