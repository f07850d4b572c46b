# tests of the wide-frame counter + ncu --set full of k_dfs after the 96-entry level B: config 4, n=40/T=500, one shard of 8 of config 5
(time timeout 900 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_sparse_midsize.py -m gpu -x -q) > gpurun_out/r2t_tests.log 2>&1; echo "sparse gpu tests rc=$?"; tail -4 gpurun_out/r2t_tests.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dfs -s 2 -c 1 -f -o gpurun_out/r2t_prof_c4 python scripts/run_sparse_once.py 24 300 0 3 > gpurun_out/r2t_ncu_c4.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dfs -s 1 -c 1 -f -o gpurun_out/r2t_prof_w2 python scripts/run_sparse_once.py 40 500 0 2 > gpurun_out/r2t_ncu_w2.log 2>&1; echo rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dfs -s 1 -c 1 -f -o gpurun_out/r2t_prof_c5s8 python scripts/run_shard_once.py 40 1000 0 8 2 > gpurun_out/r2t_ncu_c5s8.log 2>&1; echo rc=$?
tail -3 gpurun_out/r2t_ncu_c5s8.log
ls -la gpurun_out/*.ncu-rep
