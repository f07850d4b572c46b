set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc
(time python -m pytest tests -m gpu -x -q -s tests/test_gpu_sparse_midsize.py tests/test_gpu_group_kats.py) > gpurun_out/r2a_tests_new.log 2>&1; echo "new tests rc=$?"
tail -15 gpurun_out/r2a_tests_new.log
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2a_tests_all.log 2>&1; echo "all tests rc=$?"
tail -5 gpurun_out/r2a_tests_all.log
python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
cat gpurun_out/r2a_bench.json | cut -c1-3000
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err; echo "ref rc=$?"
cat gpurun_out/r2a_bench_ref.json | cut -c1-1200
python scripts/shard_sweep.py 24 300 0 8 4 > gpurun_out/r2a_shards8.log 2>&1; tail -3 gpurun_out/r2a_shards8.log
python scripts/shard_sweep.py 24 300 0 2 4 > gpurun_out/r2a_shards2.log 2>&1; tail -1 gpurun_out/r2a_shards2.log
python scripts/run_sparse_once.py 24 300 0 3 > gpurun_out/r2a_plain_c4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_dfs -s 2 -c 1 -f -o gpurun_out/r2a_prof_c4 python scripts/run_sparse_once.py 24 300 0 3 > gpurun_out/r2a_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
python scripts/run_sparse_once.py 40 500 0 3 > gpurun_out/r2a_plain_w2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_dfs -s 2 -c 1 -f -o gpurun_out/r2a_prof_w2 python scripts/run_sparse_once.py 40 500 0 3 > gpurun_out/r2a_ncu_w2.log 2>&1; echo "ncu w2 rc=$?"
cat gpurun_out/r2a_plain_c4.log gpurun_out/r2a_plain_w2.log
