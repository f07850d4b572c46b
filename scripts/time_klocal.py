"""Time-to-exact-ground-state of the DP configs (BASELINE configs 1-3) on the GPU vs the CPU oracle."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import stabilizer_gs_b200 as S
import oracle_py as O
import gen_hamiltonians as G

out = {}
def timed(f, reps=3):
    best = 1e9
    for _ in range(reps):
        t = time.perf_counter(); r = f(); best = min(best, time.perf_counter() - t)
    return r, best

ex = os.path.join(ROOT, 'examples', 'hamiltonian_finite.txt')
r, t = timed(lambda: S.klocal_search(S.HamHandle.load(ex)))
ro, to = timed(lambda: O.klocal(path=ex, threads=os.cpu_count()))
out['config1_klocal_finite'] = dict(gpu_ms=t * 1e3, cpu_oracle_ms=to * 1e3, energy=r['energy'], launches=r['kernel_launches'])
r, t = timed(lambda: S.sparse_search(S.HamHandle.load(ex)))
out['config1_sparse_finite'] = dict(gpu_ms=t * 1e3, energy=r['energy'])
pe = os.path.join(ROOT, 'examples', 'hamiltonian_periodic.txt')
r, t = timed(lambda: S.klocal_periodic(pe, cmax=10, c=3))
ro, to = timed(lambda: O.periodic(pe, cmax=10, c=3))
out['config2_periodic'] = dict(gpu_ms=t * 1e3, cpu_oracle_ms=to * 1e3, orbits=len(r['results']), min_energy_per_period=min(o['energy'] for o in r['results']) / 3,
                               python_reference_s=2.95)
for per in (3, 6):
    text = G.klocal_text(64, 4, per, 0)
    H = S.HamHandle.from_text(text)
    r, t = timed(lambda: S.klocal_search(H))
    ro, to = timed(lambda: O.klocal(text=text, threads=os.cpu_count()), reps=1)
    out[f'config3_klocal_n64_k4_per{per}'] = dict(gpu_ms=t * 1e3, gpu_device_ms=r['seconds_device'] * 1e3, cpu_oracle_ms=to * 1e3, energy=r['energy'],
                                                transitions=r['transitions'], max_states=max(r['states_per_site']), launches=r['kernel_launches'],
                                                transitions_per_s=r['transitions'] / t, python_reference_s=(116.9 if per == 3 else None))
print(json.dumps(out, indent=1))
