(time timeout 900 python -m pytest tests/test_gpu_klocal_wide.py -m gpu -x -q) > gpurun_out/r2h_tests.log 2>&1; echo "wide tests rc=$?"; tail -25 gpurun_out/r2h_tests.log
(time timeout 900 python -m pytest tests -m gpu -x -q -k "not config5 and not wide") > gpurun_out/r2h_tests_all.log 2>&1; echo "all tests rc=$?"; tail -5 gpurun_out/r2h_tests_all.log
SGS_DP_TRACE=1 python - <<'PY'
import sys, time
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import stabilizer_gs_b200 as S, gen_hamiltonians as G
for n in (64, 256, 1024):
    H = S.HamHandle.from_text(G.klocal_text(n, 4, 3, 0))
    for _ in range(2):
        t=time.time(); r = S.klocal_search(H); dt=time.time()-t
    print(n, r['energy'], len(r['generators']), r['transitions'], f"{dt*1e3:.2f} ms", flush=True)
PY
