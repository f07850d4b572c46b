timeout 300 python - <<'PY'
import sys, time, os
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
pass
import stabilizer_gs_b200 as S, gen_hamiltonians as G
for (n,k,per,seed) in [(24,6,8,3),(28,6,9,4),(32,6,10,4)]:
    H = S.HamHandle.from_text(G.klocal_text(n,k,per,seed))
    for g in (1,2):
        for _ in range(2):
            t=time.time(); r=S.klocal_search(H, ngpus=g); dt=time.time()-t
        print(n,k,per,'gpus',g, r['energy'], max(r['states_per_site']), r['transitions'], f"{dt*1e3:.1f} ms", flush=True)
PY
