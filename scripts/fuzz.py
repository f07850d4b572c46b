"""Randomised parity sweep against the CPU oracle: python scripts/fuzz.py [seconds] [seed0].
Sparse search (exhaustive, forced depths, branch and bound, virtual shards) and the k-local DP on random small
instances; prints every mismatch and exits non-zero if there was one."""
import os, sys, time, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import stabilizer_gs_b200 as S
import oracle_py as O
import gen_hamiltonians as G

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rnd = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
t0 = time.time()
bad = 0
ncase = 0


def same(a, b, keys):
    return all(a[k] == b[k] for k in keys) and abs(a['energy'] - b['energy']) <= 1e-10 * max(1.0, abs(b['energy']))


while time.time() - t0 < budget:
    seed = rnd.randrange(1 << 30)
    if rnd.random() < 0.6 or os.environ.get('FUZZ_MID'):
        if os.environ.get('FUZZ_MID'):      # mid-size: the level-B certificate / clique walk dominates
            n = rnd.choice([12, 14, 16, 18, 20, 34])
            T = rnd.randint(50, 100)
        else:
            n = rnd.choice([3, 4, 5, 6, 8, 10, 12, 14, 16, 33, 36])
            T = min(rnd.randint(4, 44 if n <= 16 else 30), 4 ** n - 1)
        text = G.sparse_text(n, T, seed)
        ref = O.sparse(text=text, threads=os.cpu_count())
        H = S.HamHandle.from_text(text)
        for kw in ({}, {'expand_depth': rnd.randint(1, 3)}, {'prune': 1}):
            r = S.sparse_search(H, **kw)
            keys = ('generators', 'term_signs') + (() if kw.get('prune') else ('candidates',))
            if not same(r, ref, keys):
                bad += 1
                print('MISMATCH sparse', n, T, seed, kw, r['energy'], ref['energy'], r['candidates'], ref['candidates'], flush=True)
        world = rnd.choice([2, 3, 5, 8])
        os.environ['SGS_SPLIT'] = '1'
        parts = [S.sparse_search(H, rank=k, world=world) for k in range(world)]
        os.environ.pop('SGS_SPLIT')
        best = min(parts, key=lambda p: p['energy'])
        if sum(p['candidates'] for p in parts) != ref['candidates'] or best['generators'] != ref['generators']:
            bad += 1
            print('MISMATCH shards', n, T, seed, world, sum(p['candidates'] for p in parts), ref['candidates'], flush=True)
    elif rnd.random() < 0.1:
        # periodic cell: l sites per period, terms of span <= k starting inside the cell
        l, k = rnd.choice([1, 2]), rnd.choice([2, 3])
        lines = []
        for st in range(l):
            for _ in range(rnd.randint(1, 2)):
                length = rnd.randint(1, k)
                letters = rnd.choice('XYZ') + ''.join(rnd.choice('IXYZ') for _ in range(length - 2)) + (rnd.choice('XYZ') if length > 1 else '')
                lines.append(f" {rnd.uniform(-1, 1):.6f} {letters} {' '.join(str(st + j) for j in range(length))}")
        text = f"{l} {k}\n" + "\n".join(sorted(set(lines))) + "\n"
        path = f"/tmp/fuzz_periodic_{os.getpid()}.txt"
        open(path, 'w').write(text)
        try:
            ref = O.periodic(path, cmax=8, c=2)
            r = S.klocal_periodic(path, cmax=8, c=2)
            key = lambda o: (round(o['energy'], 9), o['begin'], o['end'], o['generators'])
            if r['sright_log'] != ref['sright_log'] or r['fixed_point_sizes'] != ref['fixed_point_sizes'] or \
               sorted(map(key, r['results'])) != sorted(map(key, ref['results'])):
                bad += 1
                print('MISMATCH periodic', repr(text), flush=True)
        except Exception as e:       # both sides must agree on what is rejected
            print('periodic exception', repr(text), e, flush=True)
    else:
        k = rnd.randint(1, 8)
        n = rnd.randint(k, k + 14)
        per = min(rnd.randint(1, 5), 4 ** k - 1)
        text = G.klocal_text(n, k, per, seed)
        ref = O.klocal(text=text, threads=4)
        r = S.klocal_search(S.HamHandle.from_text(text))
        if not same(r, ref, ('generators', 'term_signs', 'sright_counts', 'states_per_site', 'transitions')):
            bad += 1
            print('MISMATCH klocal', n, k, per, seed, r['energy'], ref['energy'], flush=True)
    ncase += 1
print(f'fuzz: {ncase} instances, {bad} mismatches, {time.time() - t0:.1f} s')
sys.exit(1 if bad else 0)
