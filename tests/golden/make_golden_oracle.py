#!/usr/bin/env python
"""Golden vectors for instances the Python reference cannot finish (≈25 leaves/s): written by the CPU ORACLE
(oracle/sparse.c, the C restatement of FullState::evolve, cpp/state.cpp:636-687), which is itself pinned to
the unmodified Python reference by tests/test_oracle_golden.py and tests/test_oracle_live_reference.py.

    python tests/golden/make_golden_oracle.py [--only NAME ...] [--threads N]

Instances (tests/gen_hamiltonians.sparse_text(n, T, seed)):
  oracle_sparse_24_300_0   BASELINE configs[3] (the benchmark's workload): 81 206 511 leaves; ≈6 min on 8 cores
  oracle_sparse_*          mid-size instances on which the certified clique walk (level B) dominates

Output: tests/golden/<name>.json — energy, canonical generators, term signs, candidates, rank histogram,
sum of 2^rank.  The instance text is NOT stored (the generator is deterministic; its md5 is).
Nothing here is imported by the product; tests read the JSON only.
"""
import argparse
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import gen_hamiltonians as G  # noqa: E402
import oracle_py as O  # noqa: E402

INSTANCES = {
    'oracle_sparse_20_100_0': (20, 100, 0),
    'oracle_sparse_24_120_1': (24, 120, 1),
    'oracle_sparse_24_160_2': (24, 160, 2),
    'oracle_sparse_32_140_3': (32, 140, 3),
    'oracle_sparse_36_140_4': (36, 140, 4),
    'oracle_sparse_40_110_5': (40, 110, 5),
    'oracle_sparse_40_200_6': (40, 200, 6),
    'oracle_sparse_40_170_7': (40, 170, 7),     # roots of 65-96 alive terms: the three-row certificate of level B
    'oracle_sparse_30_150_8': (30, 150, 8),
    'oracle_sparse_70_150_10': (70, 150, 10),   # four-word rows (n > 64) at a size where level A / the wide level B / the hashed duplicate check run
    'oracle_sparse_100_200_11': (100, 200, 11),
    'oracle_sparse_40_260_9': (40, 260, 9),     # config 5's shape at expand_depth 1: roots of ~130 alive terms (level A), children of ~65 (three-row level B)
    'oracle_sparse_24_300_0': (24, 300, 0),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', nargs='*')
    ap.add_argument('--threads', type=int, default=os.cpu_count() or 1)
    args = ap.parse_args()
    for name, (n, T, seed) in INSTANCES.items():
        if args.only and name not in args.only:
            continue
        text = G.sparse_text(n, T, seed)
        r = O.sparse(text=text, threads=args.threads)
        out = dict(name=name, generated_by='tests/golden/make_golden_oracle.py (oracle/sparse.c, pinned to the Python reference)',
                   generator=f'gen_hamiltonians.sparse_text({n}, {T}, {seed})', md5=hashlib.md5(text.encode()).hexdigest(),
                   n=n, nterms=T, seed=seed, threads=args.threads, seconds=round(r['seconds'], 2),
                   sparse=dict(energy=r['energy'], generators=r['generators'], term_signs=r['term_signs'],
                               candidates=r['candidates'], ranks={str(k): v for k, v in r['rank_hist'].items()},
                               sum_2m=r['sum_2m']))
        with open(os.path.join(HERE, name + '.json'), 'w') as f:
            json.dump(out, f, indent=1)
        print(name, r['candidates'], r['energy'], f"{r['seconds']:.1f}s", flush=True)


if __name__ == '__main__':
    main()
