"""Mirror of the reference's py/state.py classes that its scripts touch — Hamiltonian, FullStateEnergies,
generate_Sright_all, StateMachine — as thin shims over the C ABI (all searching runs on the GPU)."""
import ctypes as C
import re

import numpy as np

from . import _capi
from .pauli import Pauli
from .stab import StabilizerGroup


class Hamiltonian:
    """py/state.py:11-72.  `.original_paulis` = [(Pauli, w)] in file order, `.paulis[qlast]` = bucket."""

    def __init__(self, n, k):
        self.n, self.k, self.l = n, k, None
        self.original_paulis = []
        self.paulis = {q: [] for q in range(n)}

    def add_Pauli(self, P, w):
        assert P.end - P.begin <= self.k
        self.paulis[P.end - 1].append((P, w))
        self.original_paulis.append((P, w))

    @classmethod
    def _parse(cls, path, periodic, cmax):
        with open(path) as f:
            a, b = map(int, f.readline().split())
            if periodic:
                H = cls(a * (1 + cmax) + 2 * b, b)
                H.l = a
            else:
                H = cls(a, b)
            for line in f:
                line = line.strip()
                if not line:
                    break                                  # py/state.py:41-43
                args = re.split(r'\s+', line)
                w, string, qubits = float(args[0]), args[1], list(map(int, args[2:]))
                if periodic:
                    for i in range(-H.n, H.n + 1):
                        if max(qubits) + i * a < H.n and min(qubits) + i * a >= 0:
                            H.add_Pauli(Pauli.from_string(string, [q + i * a for q in qubits]), w)
                else:
                    H.add_Pauli(Pauli.from_string(string, qubits), w)
        H.path, H.periodic, H.cmax = path, periodic, cmax
        return H

    @classmethod
    def from_file(cls, path):
        return cls._parse(path, False, 0)

    @classmethod
    def from_file_periodic(cls, path, cmax):
        return cls._parse(path, True, cmax)

    def handle(self):
        """sgs_ham* built from the in-memory terms (sgs_ham_from_arrays)."""
        T, n = len(self.original_paulis), self.n
        W = (2 * n + 63) // 64
        zx = (C.c_uint64 * max(T * W, 1))()
        w = (C.c_double * max(T, 1))()
        b = (C.c_int * max(T, 1))()
        e = (C.c_int * max(T, 1))()
        for t, (P, wt) in enumerate(self.original_paulis):
            row = Pauli(P.z[None], P.x[None], P.begin).packed(n)[0]
            for k in range(W):
                zx[t * W + k] = row[k]
            w[t], b[t], e[t] = wt, P.begin, P.end
        p = C.c_void_p()
        _capi.check(_capi.lib().sgs_ham_from_arrays(n, self.k, T, zx, w, b, e, C.byref(p)))
        return _capi.HamHandle(p)


def _group_from_result(r, n):
    rows = []
    for s in r['generators']:
        words = [0] * ((2 * n + 63) // 64)
        for j, ch in enumerate(s[1:]):
            v = {'I': 0, 'Z': 1, 'X': 2, 'Y': 3}[ch]
            words[(2 * j) >> 6] |= v << ((2 * j) & 63)
        rows.append(words)
    signs = [1 if s[0] == '+' else -1 for s in r['generators']]
    return Pauli.from_packed(rows, signs, n)


class FullStateEnergies:
    """py/state.py:88-147: `.evolve()` → (signed generators as a Pauli, ground-state energy)."""

    def __init__(self, paulis, **opts):
        n = max(P.end for P, w in paulis)
        k = max(P.end - P.begin for P, w in paulis)
        self.H = Hamiltonian(n, max(k, 1))
        for P, w in paulis:
            self.H.add_Pauli(P, w)
        self.opts = opts

    def evolve(self):
        r = _capi.sparse_search(self.H.handle(), **self.opts)
        self.result = r
        print('number of states', r['candidates'])
        return _group_from_result(r, self.H.n), r['energy']


class _SrightAll:
    """what generate_Sright_all returns: the right-environment tables live inside the sgs_klocal handle"""

    def __init__(self, H, ptr, ham):
        self.H, self.ptr, self.ham = H, ptr, ham

    def __del__(self):
        if getattr(self, 'ptr', None) and _capi._lib is not None:
            _capi._lib.sgs_klocal_free(self.ptr)
            self.ptr = None


def generate_Sright_all(H, **opts):
    """py/state.py:526-539 (prints "m count" per site, last site first)."""
    o, keep = _capi.make_opts(**opts)
    ham = H.handle()
    p = C.c_void_p()
    _capi.check(_capi.lib().sgs_klocal_create(ham.ptr, C.byref(o), -1, C.byref(p)))
    for m in range(H.n - 1, -1, -1):
        print(m, _capi.lib().sgs_klocal_sright_count(p, m))
    return _SrightAll(H, p, ham)


class _StateDict:
    def __init__(self, n):
        self._n = n

    def __len__(self):
        return self._n


class StateMachine:
    """py/state.py:328-399: `.evolve()`, `.m`, `len(.state_dict)`, `.generate_gs()` → `.energy_gs`, `.stab_gs`."""

    def __init__(self, hamiltonian, Sright_all):
        self.hamiltonian, self.n, self.k = hamiltonian, hamiltonian.n, hamiltonian.k
        self._sr = Sright_all

    @property
    def m(self):
        return _capi.lib().sgs_klocal_m(self._sr.ptr)

    @property
    def state_dict(self):
        return _StateDict(_capi.lib().sgs_klocal_nstates(self._sr.ptr))

    def evolve(self):
        _capi.check(_capi.lib().sgs_klocal_evolve(self._sr.ptr))

    def generate_gs(self):
        r = _capi.Result()
        _capi.check(_capi.lib().sgs_klocal_ground_state(self._sr.ptr, C.byref(r)))
        try:
            d = _capi.result_to_dict(r, self.n)
        finally:
            _capi.lib().sgs_result_free(C.byref(r))
        self.energy_gs = d['energy']
        self.stab_gs = StabilizerGroup(_group_from_result(d, self.n), canonicalized=True)
        self.term_signs = d['term_signs']
