timeout 280 python -m pytest tests/test_gpu_sparse_midsize.py -m gpu -x -q -k "rank_limit" > gpurun_out/r2b_ranklimit.log 2>&1; echo "ranklimit rc=$?"; tail -5 gpurun_out/r2b_ranklimit.log
timeout 600 scripts/ab_libs.sh build_ab/libsgs_base.so build_ab/libsgs_m6.so build_ab/libsgs_m8.so build_ab/libsgs_m10.so build_ab/libsgs_w8m4.so 2>&1 | tee gpurun_out/r2b_ab.log
timeout 120 python scripts/run_sparse_once.py 40 1000 0 1 2>&1 | tee gpurun_out/r2b_c5.log
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r2b_tests.log
