# single-pass level A where the child list fits whatever its length; O(n) duplicate check for lists of 33-128 entries
(time timeout 900 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_sparse_midsize.py tests/test_gpu_limits.py -m gpu -x -q) > gpurun_out/r2v_tests.log 2>&1; echo "sparse gpu tests rc=$?"; tail -4 gpurun_out/r2v_tests.log
for a in "24 300 0 4" "30 400 0 3" "40 500 0 3" "40 1000 0 1"; do echo "== $a"; timeout 300 python scripts/run_sparse_once.py $a | tail -1 | sed 's/.*slow=/slow=/'; done
timeout 120 python scripts/e2e_trace.py 2>&1 | grep -v "level\|work stack" | tail -12
