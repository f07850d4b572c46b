timeout 300 python - <<'PY' 2>&1 | grep -v "NCCL version" | tail -12
import sys, time, os
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
os.environ['SGS_DP_TRACE']='1'
import stabilizer_gs_b200 as S, gen_hamiltonians as G
H = S.HamHandle.from_text(G.klocal_text(28,6,9,4))
for g in (1,2):
    t=time.time(); r=S.klocal_search(H, ngpus=g); dt=time.time()-t
    print('gpus',g, r['energy'], max(r['states_per_site']), r['transitions'], f"{dt*1e3:.1f} ms", flush=True)
PY
