# config 5: which paths does the walk take, and where is the time?
cp stabilizer_gs_b200/libsgs.so /tmp/libsgs_saved.so
cp build_ab/libsgs_trace.so stabilizer_gs_b200/libsgs.so
SGS_PERM=1 SGS_TRACE=1 timeout 300 python - <<'PY' 2>&1 | grep -v "work stack"
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import stabilizer_gs_b200 as S, gen_hamiltonians as G
for n,T in ((40,500),(40,1000)):
    plan = S.SparsePlan(S.HamHandle.from_text(G.sparse_text(n,T,0)))
    for _ in range(2):
        r = plan.run()
        print(n,T,'dev %.3f ms hot %.3f'%(r['seconds_device']*1e3, r['seconds_hot_kernel']*1e3), 'paths', r['walk_paths'], 'hist', dict(list(r['rank_hist'].items())[:16]) if isinstance(r['rank_hist'], dict) else r['rank_hist'], flush=True)
    plan.close()
PY
cp /tmp/libsgs_saved.so stabilizer_gs_b200/libsgs.so
