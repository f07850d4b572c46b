(time timeout 900 python -m pytest tests -m gpu -x -q -k "not config5") > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2c_tests.log
timeout 600 scripts/ab_libs.sh build_ab/libsgs_base.so build_ab/libsgs_m6.so build_ab/libsgs_m8.so 2>&1 | tee gpurun_out/r2c_ab.log
timeout 120 python scripts/run_sparse_once.py 40 1000 0 1 2>&1 | tee gpurun_out/r2c_c5.log
python scripts/run_sparse_once.py 24 300 0 3 > gpurun_out/r2c_plain_c4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_dfs -s 2 -c 1 -f -o gpurun_out/r2c_prof_c4 python scripts/run_sparse_once.py 24 300 0 3 > gpurun_out/r2c_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
