# k_finalize: ordered member compaction (no serial sort); are the hand-off rounds of the sharded walk still worth it?
(time timeout 900 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_sparse_midsize.py tests/test_gpu_limits.py tests/test_gpu_python_api.py -m gpu -x -q) > gpurun_out/r2p_tests.log 2>&1; echo "sparse gpu tests rc=$?"; tail -4 gpurun_out/r2p_tests.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 16 --csv --log-file gpurun_out/r2p_launches.csv python scripts/run_sparse_once.py 24 300 0 2 > gpurun_out/r2p_ncu1.log 2>&1; echo "launch list rc=$?"; grep -c k_ gpurun_out/r2p_launches.csv
for w in 8 4; do
  echo "== config 4, world $w: default"; timeout 300 python scripts/shard_sweep.py 24 300 0 $w 6 2>&1 | tail -1
  echo "== config 4, world $w: SGS_NOSPLIT"; SGS_NOSPLIT=1 timeout 300 python scripts/shard_sweep.py 24 300 0 $w 6 2>&1 | tail -1
  echo "== config 4, world $w: SGS_SPLIT"; SGS_SPLIT=1 timeout 300 python scripts/shard_sweep.py 24 300 0 $w 6 2>&1 | tail -1
done
echo "== n=40 T=500, world 8: default"; timeout 300 python scripts/shard_sweep.py 40 500 0 8 3 2>&1 | tail -1
echo "== n=40 T=500, world 8: SGS_NOSPLIT"; SGS_NOSPLIT=1 timeout 300 python scripts/shard_sweep.py 40 500 0 8 3 2>&1 | tail -1
timeout 600 python bench.py --no-config5 > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2p_bench.json')); print(d['ms_per_step'], d['roofline']['hot_kernel_ms'], d['e2e']['ms_per_step'])"
