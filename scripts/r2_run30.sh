# lists without weights (gathered per level-B frame), arena fitted to the occupancy steps
(time timeout 900 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_sparse_midsize.py tests/test_gpu_limits.py -m gpu -x -q) > gpurun_out/r2u_tests.log 2>&1; echo "sparse gpu tests rc=$?"; tail -4 gpurun_out/r2u_tests.log
for a in "24 300 0 4" "30 400 0 3" "40 500 0 3" "40 1000 0 1"; do echo "== $a"; timeout 300 python scripts/run_sparse_once.py $a | tail -1 | sed 's/.*slow=/slow=/'; done
echo "== prune"; timeout 300 python scripts/prune_time.py 40 500 0 2>&1 | tail -3
