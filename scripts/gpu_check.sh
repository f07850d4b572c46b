# full GPU check of the tree: every -m gpu test, smoke(), the default bench line and the reference arm
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/chk_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/chk_tests.log
(time timeout 300 python __graft_entry__.py smoke) > gpurun_out/chk_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/chk_smoke.log
(time timeout 900 python bench.py) > gpurun_out/chk_bench.json 2> gpurun_out/chk_bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/chk_bench.json
(time timeout 600 python bench.py --impl reference) > gpurun_out/chk_bench_ref.json 2> gpurun_out/chk_bench_ref.err; echo "ref rc=$?"; tail -c 1500 gpurun_out/chk_bench_ref.json
(time timeout 400 python scripts/fuzz.py 45 20261021) > gpurun_out/chk_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -3 gpurun_out/chk_fuzz.log
