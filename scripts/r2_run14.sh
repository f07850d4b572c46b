nvidia-smi -L | wc -l
(time timeout 600 python -m pytest tests/test_gpu_multi_gpu.py -m gpu -x -q) > gpurun_out/r2m_tests.log 2>&1; echo "multi tests rc=$?"; tail -25 gpurun_out/r2m_tests.log
timeout 300 python - <<'PY'
import sys, time
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import stabilizer_gs_b200 as S, gen_hamiltonians as G
for (n,k,per,seed) in [(64,4,6,0),(24,6,8,3)]:
    H = S.HamHandle.from_text(G.klocal_text(n,k,per,seed))
    for g in (1,2):
        for _ in range(2):
            t=time.time(); r=S.klocal_search(H, ngpus=g); dt=time.time()-t
        print(n,k,per,'gpus',g, r['energy'], max(r['states_per_site']), r['transitions'], f"{dt*1e3:.1f} ms", flush=True)
PY
