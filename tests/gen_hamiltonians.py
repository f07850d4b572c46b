"""Synthetic Hamiltonian generators (data spec from SURVEY.md Appendix C).

These reproduce, byte for byte, the instance files whose md5s are listed in
SURVEY.md Appendix C / BASELINE.md (NumPy `default_rng` = PCG64).  Letter code
0:I 1:X 2:Z 3:Y is the reference's `array4 = 2z + x` (cpp/stab.cpp:201-203,359).
"""
import numpy as np


def sparse_text(n, T, seed):
    """BASELINE configs 4/5: T distinct non-identity uniform Paulis on n qubits."""
    rng = np.random.default_rng(seed)
    seen = set()
    lines = []
    while len(lines) < T:
        a = rng.integers(0, 4, size=n)
        if not a.any():
            continue
        key = tuple(a)
        if key in seen:
            continue
        seen.add(key)
        sites = [i for i in range(n) if a[i]]
        s = ''.join('IXZY'[a[i]] for i in sites)
        w = rng.uniform(-1, 1)
        lines.append(f"{w: .10f} {s} {' '.join(map(str, sites))}")
    return f"{n} {n}\n" + '\n'.join(lines) + '\n'


def klocal_text(n, k, per_site, seed):
    """BASELINE config 3: per_site distinct random Paulis in every k-site window."""
    rng = np.random.default_rng(seed)
    seen = set()
    lines = []
    for m in range(n - k + 1):
        c = 0
        while c < per_site:
            a = rng.integers(0, 4, size=k)
            if not a.any():
                continue
            sites = [m + i for i in range(k) if a[i]]
            s = ''.join('IXZY'[a[i]] for i in range(k) if a[i])
            key = (s, tuple(sites))
            if key in seen:
                continue
            seen.add(key)
            c += 1
            lines.append(f"{rng.uniform(-1, 1): .10f} {s} {' '.join(map(str, sites))}")
    return f"{n} {k}\n" + '\n'.join(lines) + '\n'


def gen_sparse(n, T, seed, path):
    with open(path, 'w') as f:
        f.write(sparse_text(n, T, seed))


def gen_klocal(n, k, per_site, seed, path):
    with open(path, 'w') as f:
        f.write(klocal_text(n, k, per_site, seed))
