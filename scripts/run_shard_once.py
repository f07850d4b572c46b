"""Time one virtual shard: python scripts/run_shard_once.py n T seed world [reps]  (rank 0 of `world`, no collective)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import stabilizer_gs_b200 as S
import gen_hamiltonians as G
n, T, seed, world = map(int, sys.argv[1:5])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 4
for rank in (0, world - 1):
    plan = S.SparsePlan(S.HamHandle.from_text(G.sparse_text(n, T, seed)), rank=rank, world=world)
    for _ in range(reps):
        r = plan.run()
    print(f"world={world} rank={rank} cand={r['candidates']} dev={r['seconds_device']*1e3:.3f}ms hot={r['seconds_hot_kernel']*1e3:.3f}ms", flush=True)
    plan.close()
