(time timeout 900 python -m pytest tests -m gpu -x -q -k "not config5") > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2d_tests.log
for i in 1 2; do python scripts/run_sparse_once.py 24 300 0 5 | tail -1; python scripts/run_sparse_once.py 40 500 0 3 | tail -1; done
SGS_TRACE=1 python scripts/shard_sweep.py 24 300 0 8 3 > gpurun_out/r2d_shards8.log 2>&1; grep -A3 "rank=0\|replay shard 0" gpurun_out/r2d_shards8.log | tail -12; tail -1 gpurun_out/r2d_shards8.log
SGS_TRACE=1 python scripts/run_sparse_once.py 24 300 0 3 2>&1 | tail -6
python scripts/run_sparse_once.py 24 300 0 3 > gpurun_out/r2d_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2d_launches.csv python scripts/run_sparse_once.py 24 300 0 3 > gpurun_out/r2d_ncu.log 2>&1; echo "ncu rc=$?"
