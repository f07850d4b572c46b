# hand-off rounds chosen by roots per resident warp (host rule)
(time timeout 900 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_sparse_midsize.py tests/test_gpu_limits.py tests/test_gpu_python_api.py -m gpu -x -q) > gpurun_out/r2q_tests.log 2>&1; echo "sparse gpu tests rc=$?"; tail -4 gpurun_out/r2q_tests.log
run() { echo "== $*"; env "${@:5}" timeout 300 python scripts/shard_sweep.py $1 $2 0 $3 $4 2>&1 | tail -1 | sed 's/.*| //'; }
run 24 300 8 6 A=1
run 24 300 4 6 A=1
run 24 300 3 6 A=1
run 24 300 2 6 A=1
run 40 500 8 3 A=1
run 40 500 2 3 A=1
run 30 400 8 3 A=1
run 30 400 4 3 A=1
run 30 400 4 3 SGS_SPLIT=1
run 20 200 8 6 A=1
run 20 200 8 6 SGS_NOSPLIT=1
timeout 600 python bench.py --no-config5 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2q_bench.json')); print(d['ms_per_step'], d['roofline']['hot_kernel_ms'], d['e2e']['ms_per_step'])"
