"""First-contact GPU script: run many sparse instances and print (not assert) every mismatch vs the oracle."""
import json, os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import stabilizer_gs_b200 as S
import oracle_py as O
import gen_hamiltonians as G

def cmp(tag, text, depth):
    ref = O.sparse(text=text, threads=4)
    try:
        t0 = time.time()
        r = S.sparse_search(S.HamHandle.from_text(text), expand_depth=depth)
        dt = time.time() - t0
    except Exception as e:
        print(f"[ERR ] {tag} depth={depth}: {e}"); return False
    bad = []
    if abs(r['energy'] - ref['energy']) > 1e-10 * max(1, abs(ref['energy'])): bad.append(f"E {r['energy']} vs {ref['energy']}")
    if r['generators'] != ref['generators']: bad.append(f"gens {r['generators']} vs {ref['generators']}")
    if r['term_signs'] != ref['term_signs']: bad.append("term_signs")
    if r['candidates'] != ref['candidates']: bad.append(f"cand {r['candidates']} vs {ref['candidates']}")
    if {int(k): v for k, v in r['rank_hist'].items()} != ref['rank_hist']: bad.append(f"hist {r['rank_hist']} vs {ref['rank_hist']}")
    print(f"[{'OK  ' if not bad else 'FAIL'}] {tag} depth={depth} cand={r['candidates']} slow={r['slow_nodes']} launches={r['kernel_launches']} t={dt*1e3:.1f}ms {'; '.join(bad)}", flush=True)
    return not bad

print(S._capi.lib().sgs_version().decode(), 'devices', S._capi.lib().sgs_device_count(), flush=True)
ok = True
ex = open(os.path.join(ROOT, 'examples/hamiltonian_finite.txt')).read()
for d in (0, 1, 2, 3):
    ok &= cmp('finite_example', ex, d)
for (n, T, seed) in [(3, 14, 5), (4, 16, 7), (6, 12, 1), (6, 20, 1), (8, 30, 1), (10, 30, 1), (12, 36, 8), (16, 40, 13), (40, 16, 0)]:
    for d in (0, 1, 2):
        ok &= cmp(f'sparse({n},{T},{seed})', G.sparse_text(n, T, seed), d)
for (n, k, p, seed) in [(10, 3, 2, 1), (9, 4, 2, 2)]:
    for d in (0, 2):
        ok &= cmp(f'klocal({n},{k},{p},{seed})', G.klocal_text(n, k, p, seed), d)
print('ALL OK' if ok else 'SOME FAILED', flush=True)
# throughput peek
for (n, T, seed) in [(16, 80, 0), (20, 150, 0), (24, 300, 0)]:
    try:
        r = S.sparse_search(S.HamHandle.from_text(G.sparse_text(n, T, seed)))
        print(f"n={n} T={T}: E={r['energy']:.10f} m={r['m']} cand={r['candidates']:.4g} slow={r['slow_nodes']} dev={r['seconds_device']*1e3:.2f}ms hot={r['seconds_hot_kernel']*1e3:.2f}ms total={r['seconds_total']*1e3:.2f}ms launches={r['kernel_launches']} rate={r['candidates']/max(r['seconds_device'],1e-9):.3g}/s hist={r['rank_hist']}", flush=True)
    except Exception as e:
        print('ERR', n, T, e, flush=True)
print(S.measure_int_peaks(0))
