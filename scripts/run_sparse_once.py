"""Run one sparse search (for profiling): python scripts/run_sparse_once.py n T seed [repeats] [expand_depth]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import stabilizer_gs_b200 as S
import gen_hamiltonians as G
n, T, seed = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
depth = int(sys.argv[5]) if len(sys.argv) > 5 else 0
plan = S.SparsePlan(S.HamHandle.from_text(G.sparse_text(n, T, seed)), expand_depth=depth)
for _ in range(reps):
    r = plan.run()
    print(f"n={n} T={T} E={r['energy']:.10f} m={r['m']} cand={r['candidates']} slow={r['slow_nodes']} dev={r['seconds_device']*1e3:.3f}ms hot={r['seconds_hot_kernel']*1e3:.3f}ms total={r['seconds_total']*1e3:.3f}ms rate={r['candidates']/r['seconds_device']:.4g}/s", flush=True)
plan.close()
