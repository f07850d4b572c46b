# level B on up to 96 entries (<= 64 candidates), heavy-first order always
(time timeout 900 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_sparse_midsize.py -m gpu -x -q) > gpurun_out/r2s_tests.log 2>&1; echo "sparse gpu tests rc=$?"; tail -4 gpurun_out/r2s_tests.log
for a in "24 300 0 4" "30 400 0 3" "40 500 0 3" "32 140 0 3" "40 1000 0 1"; do echo "== $a"; timeout 300 python scripts/run_sparse_once.py $a | tail -1 | sed 's/.*slow=/slow=/'; done
