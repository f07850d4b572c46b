set -x
python bench.py > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"
cut -c1-1500 gpurun_out/r2k_bench.json
python scripts/run_sparse_once.py 24 300 0 3 > gpurun_out/r2k_plain_c4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_dfs -s 2 -c 1 -f -o gpurun_out/r2k_prof_c4 python scripts/run_sparse_once.py 24 300 0 3 > gpurun_out/r2k_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
python scripts/run_sparse_once.py 40 500 0 3 > gpurun_out/r2k_plain_w2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_dfs -s 2 -c 1 -f -o gpurun_out/r2k_prof_w2 python scripts/run_sparse_once.py 40 500 0 3 > gpurun_out/r2k_ncu_w2.log 2>&1; echo "ncu w2 rc=$?"
python scripts/run_sparse_once.py 40 1000 0 1 > gpurun_out/r2k_plain_c5.log 2>&1 && \
ncu --metrics smsp__inst_executed.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.per_cycle_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,launch__registers_per_thread --clock-control none -k regex:k_dfs -c 1 -f -o gpurun_out/r2k_prof_c5 python scripts/run_sparse_once.py 40 1000 0 1 > gpurun_out/r2k_ncu_c5.log 2>&1; echo "ncu c5 rc=$?"
python scripts/run_sparse_once.py 24 300 0 3 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2k_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-config5 > gpurun_out/r2k_ncu_bench.log 2>&1; echo "ncu bench rc=$?"
cat gpurun_out/r2k_plain_c4.log gpurun_out/r2k_plain_w2.log gpurun_out/r2k_plain_c5.log
