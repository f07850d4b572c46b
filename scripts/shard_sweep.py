"""Emulate an N-GPU step on ONE GPU: every virtual shard of `world` runs alone, one after the other (no collective);
the slowest shard bounds the N-GPU device time.  python scripts/shard_sweep.py n T seed world [reps]
SGS_TRACE=1 prints the phases of each replayed step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import stabilizer_gs_b200 as S
import gen_hamiltonians as G
n, T, seed, world = map(int, sys.argv[1:5])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 4
H = S.HamHandle.from_text(G.sparse_text(n, T, seed))
one = S.SparsePlan(H)
for _ in range(reps):
    r1 = one.run()
one.close()
devs, hots, cands = [], [], []
for rank in range(world):
    plan = S.SparsePlan(H, rank=rank, world=world)
    best = None
    for _ in range(reps):
        r = plan.run()
        if best is None or r['seconds_device'] < best['seconds_device']:
            best = r
    plan.close()
    devs.append(best['seconds_device'] * 1e3); hots.append(best['seconds_hot_kernel'] * 1e3); cands.append(best['candidates'])
    print(f"world={world} rank={rank} cand={best['candidates']} dev={devs[-1]:.3f}ms hot={hots[-1]:.3f}ms launches={best['kernel_launches']}", flush=True)
assert sum(cands) == r1['candidates']
print(f"n={n} T={T} one shard: dev={r1['seconds_device']*1e3:.3f}ms hot={r1['seconds_hot_kernel']*1e3:.3f}ms | world={world}: max dev={max(devs):.3f}ms "
      f"max hot={max(hots):.3f}ms mean hot={sum(hots)/world:.3f}ms cand imbalance={max(cands)*world/sum(cands):.3f} "
      f"=> speed-up {r1['seconds_device']*1e3/max(devs):.2f}x (no collective), walk-only {r1['seconds_hot_kernel']*1e3/max(hots):.2f}x")
