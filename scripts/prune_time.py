"""exhaustive vs branch-and-bound time-to-exact-ground-state: python scripts/prune_time.py n T seed"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import stabilizer_gs_b200 as S
import gen_hamiltonians as G
n, T, seed = map(int, sys.argv[1:4])
H = S.HamHandle.from_text(G.sparse_text(n, T, seed))
skip_full = len(sys.argv) > 4
for prune in ((1,) if skip_full else (0, 1)):
    plan = S.SparsePlan(H, prune=prune)
    for _ in range(3):
        t = time.perf_counter(); r = plan.run(); w = time.perf_counter() - t
    plan.close()
    print(f"n={n} T={T} prune={prune} E={r['energy']:.10f} m={r['m']} visited={r['candidates']} dev={r['seconds_device']*1e3:.3f}ms wall={w*1e3:.3f}ms gens0={r['generators'][0]}", flush=True)
