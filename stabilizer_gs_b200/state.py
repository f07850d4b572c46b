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
        with open(path) as f:
            H.text = f.read()
        return H

    @classmethod
    def from_file(cls, path):
        return cls._parse(path, False, 0)

    @classmethod
    def from_file_periodic(cls, path, cmax):
        return cls._parse(path, True, cmax)

    def handle(self):
        """sgs_ham* built from the in-memory terms (sgs_ham_from_arrays); a periodic Hamiltonian is re-parsed
        from its file text (the handle has to know the period)."""
        if getattr(self, 'periodic', False):
            return _capi.HamHandle.from_text(self.text, periodic=True, cmax=self.cmax)
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


def _idx_flat(index, rank):
    """sign-pattern tuple (generator 0 first) -> flat table index: bit (rank-1-i) set <=> generator i is '-'
    (py/stab.py:176-190 keeps the table as an ndarray of shape (2,)*rank: the same C-order index)"""
    flat = 0
    for i, b in enumerate(index):
        flat |= int(b) << (rank - 1 - i)
    return flat


class EnergyLevel:
    """py/stab.py:176-305, the part the scripts touch: `.Es` (ndarray of shape (2,)*rank, assignable), `.clear()`,
    `.get_state(index, paulis, return_stab)`."""

    def __init__(self, state):
        self._state = state

    @property
    def Es(self):
        st = self._state
        buf = (C.c_double * (1 << st.rank))()
        _capi.check(_capi.lib().sgs_klocal_state_energies(st._lst, 0, buf))
        return np.array(buf[:], dtype=float).reshape((2,) * st.rank)

    @Es.setter
    def Es(self, value):
        st = self._state
        flat = np.ascontiguousarray(np.asarray(value, dtype=float).reshape(-1))
        assert flat.size == 1 << st.rank
        _capi.check(_capi.lib().sgs_klocal_state_set_energies(st._lst, 0, flat.ctypes.data_as(C.POINTER(C.c_double))))

    def clear(self):                                           # py/stab.py:282-286
        self.Es = np.zeros((2,) * self._state.rank)
        self._state._hist = {}
        self._state._hist_live = False

    def get_state(self, index, paulis, return_stab=True):      # py/stab.py:288-305
        st = self._state
        flat = _idx_flat(index, st.rank)
        terms, signs = st._history(flat)
        energy = float(self.Es.reshape(-1)[flat])
        stack = Pauli.stack([paulis[t][0] for t in terms]) * np.array(signs) if terms else Pauli(np.zeros((0, 0)), np.zeros((0, 0)))
        if return_stab:
            return StabilizerGroup(stack), energy
        return stack, energy


class _WindowGroup:
    """State.stab as the scripts see it: the window group (all generators '+') with its energy table"""

    def __init__(self, state):
        self._state = state
        self.energy_level = EnergyLevel(state)

    @property
    def begin(self):
        return self._state._info()[2]

    @property
    def end(self):
        return self._state._info()[3]

    @property
    def Ps(self):
        st = self._state
        n = st._mach.H.n
        W = (2 * n + 63) // 64
        buf = (C.c_uint64 * max(st.rank * W, 1))()
        _capi.check(_capi.lib().sgs_klocal_state_rows(st._mach.ptr, st._lst, 0, buf))
        rows = [[buf[i * W + k] for k in range(W)] for i in range(st.rank)]
        return Pauli.from_packed(rows, [1] * st.rank, n, self.begin, self.end)

    def __repr__(self):
        return repr(self.Ps)


class State:
    """py/state.py:150-325 as a handle on one DP state held by libsgs: `.m`, `.stab.energy_level.Es`, `hash()`,
    `.copy()`, `.shift(add, clear_history)`.  The state's content (window group, valid terms, right environment,
    energy table) lives in a one-entry sgs_klocal_states list; the per-entry history (the reference's
    Ps_indices / Ps_signs) is walked on the device and cached here when a shift keeps it."""

    def __init__(self, mach, lst, hist=None, hist_live=False):
        self._mach, self._lst = mach, lst
        self._hist = dict(hist or {})
        self._hist_live = hist_live          # the device still holds this state's history
        self.stab = _WindowGroup(self)

    def __del__(self):
        if getattr(self, '_lst', None) and _capi._lib is not None:
            _capi._lib.sgs_klocal_states_free(self._lst)
            self._lst = None

    def _info(self):
        v = [C.c_int() for _ in range(5)]
        _capi.check(_capi.lib().sgs_klocal_state_info(self._mach.ptr, self._lst, 0, *[C.byref(x) for x in v]))
        return [x.value for x in v]          # m, rank, begin, end, sright

    @property
    def m(self):
        return self._info()[0]

    @property
    def rank(self):
        return self._info()[1]

    def _key(self):
        key = (C.c_uint64 * 32)()
        nw = C.c_int()
        _capi.check(_capi.lib().sgs_klocal_state_key(self._mach.ptr, self._lst, 0, key, C.byref(nw)))
        return tuple(key[:nw.value])

    def __hash__(self):                                        # py/state.py:224-225 (energies and history are not in the key)
        return hash(self._key())

    def __eq__(self, other):
        return isinstance(other, State) and self._key() == other._key()

    def _history(self, flat):
        if flat in self._hist:
            return self._hist[flat]
        if not self._hist_live:
            return [], []
        cap = 4 * (self._mach.H.n + 4) * max(self._mach.H.k, 1) + 64
        terms, signs, cnt = (C.c_int * cap)(), (C.c_int8 * cap)(), C.c_int()
        _capi.check(_capi.lib().sgs_klocal_state_history(self._mach.ptr, self._lst, 0, flat, terms, signs, cap, C.byref(cnt)))
        self._hist[flat] = (list(terms[:cnt.value]), [int(x) for x in signs[:cnt.value]])
        return self._hist[flat]

    def copy(self):
        idx = (C.c_int * 1)(0)
        p = C.c_void_p()
        _capi.check(_capi.lib().sgs_klocal_states_select(self._lst, idx, 1, C.byref(p)))
        return State(self._mach, p, self._hist, self._hist_live)

    def shift(self, add, clear_history=False):                 # py/state.py:309-322
        if not clear_history and self._hist_live:
            for flat in range(1 << self.rank):                 # the machine will move on: fetch the history now
                self._history(flat)
        new = self.copy()
        _capi.check(_capi.lib().sgs_klocal_states_shift(self._mach.ptr, new._lst, add, 1 if clear_history else 0))
        new._hist_live = False
        if clear_history:
            new._hist = {}
        return new

    def clear_history(self):                                   # py/state.py:324-325
        self.stab.energy_level.clear()


class _Machine:
    """what generate_Sright_all / generate_Sright_periodic return: the right-environment tables (and the DP
    machine built on them) live inside one sgs_klocal handle"""

    def __init__(self, H, ptr, ham):
        self.H, self.ptr, self.ham = H, ptr, ham

    def __del__(self):
        if getattr(self, 'ptr', None) and _capi._lib is not None:
            _capi._lib.sgs_klocal_free(self.ptr)
            self.ptr = None

    def states(self):
        """list(state_dict.values()) of the current site, key order"""
        L = _capi.lib()
        p = C.c_void_p()
        _capi.check(L.sgs_klocal_get_states(self.ptr, C.byref(p)))
        try:
            out = []
            for i in range(L.sgs_klocal_states_count(p)):
                idx = (C.c_int * 1)(i)
                q = C.c_void_p()
                _capi.check(L.sgs_klocal_states_select(p, idx, 1, C.byref(q)))
                out.append(State(self, q, hist_live=True))
            return out
        finally:
            L.sgs_klocal_states_free(p)


_SrightAll = _Machine


def generate_Sright_all(H, **opts):
    """py/state.py:526-539 (prints "m count" per site, last site first)."""
    o, keep = _capi.make_opts(**opts)
    ham = H.handle()
    p = C.c_void_p()
    _capi.check(_capi.lib().sgs_klocal_create(ham.ptr, C.byref(o), -1, C.byref(p)))
    for m in range(H.n - 1, -1, -1):
        print(m, _capi.lib().sgs_klocal_sright_count(p, m))
    return _Machine(H, p, ham)


def generate_Sright_periodic(H, cmax, **opts):
    """py/state.py:542-567: right environments of the unrolled periodic chain, iterated over one period until
    they reach a fixed point (prints the same "m count" lines)."""
    o, keep = _capi.make_opts(**opts)
    ham = H.handle()
    p = C.c_void_p()
    _capi.check(_capi.lib().sgs_klocal_create(ham.ptr, C.byref(o), cmax, C.byref(p)))
    L = _capi.lib()
    npairs = L.sgs_klocal_sright_log(p, None, 0)
    buf = (C.c_int * max(2 * npairs, 1))()
    L.sgs_klocal_sright_log(p, buf, npairs)
    for i in range(npairs):
        print(buf[2 * i], buf[2 * i + 1])
    return _Machine(H, p, ham)


class StateMachine:
    """py/state.py:328-399: `.evolve()`, `.m`, `.state_dict` ({hash(state): State}, assignable), `.generate_gs()` →
    `.energy_gs`, `.stab_gs`."""

    def __init__(self, hamiltonian, Sright_all):
        self.hamiltonian, self.n, self.k = hamiltonian, hamiltonian.n, hamiltonian.k
        self._sr = Sright_all
        self._pending, self._pending_m, self._cache = None, None, None

    @property
    def m(self):
        return self._pending_m if self._pending_m is not None else _capi.lib().sgs_klocal_m(self._sr.ptr)

    @m.setter
    def m(self, value):
        self._pending_m = int(value)
        if self._pending is None:                     # only the site changes: keep the machine's states
            self._pending = list(self.state_dict.values())

    @property
    def state_dict(self):
        if self._pending is not None:
            return {hash(st): st for st in self._pending}
        if self._cache is None:
            self._cache = {hash(st): st for st in self._sr.states()}
        return self._cache

    @state_dict.setter
    def state_dict(self, value):
        self._pending = list(value.values())
        self._cache = None

    def _flush(self):
        if self._pending is None:
            return
        L = _capi.lib()
        states, m = self._pending, self._pending_m
        self._pending, self._pending_m = None, None
        if not states:
            raise ValueError('state_dict is empty')
        idx = (C.c_int * 1)(0)
        lst = C.c_void_p()
        _capi.check(L.sgs_klocal_states_select(states[0]._lst, idx, 0, C.byref(lst)))      # an empty list
        try:
            for st in states:
                _capi.check(L.sgs_klocal_states_append(lst, st._lst, 0))
            _capi.check(L.sgs_klocal_set_states(self._sr.ptr, lst, states[0].m if m is None else m))
        finally:
            L.sgs_klocal_states_free(lst)

    def evolve(self):
        self._flush()
        self._cache = None
        _capi.check(_capi.lib().sgs_klocal_evolve(self._sr.ptr))

    def generate_gs(self):
        self._flush()
        r = _capi.Result()
        _capi.check(_capi.lib().sgs_klocal_ground_state(self._sr.ptr, C.byref(r)))
        try:
            d = _capi.result_to_dict(r, self.n)
        finally:
            _capi.lib().sgs_result_free(C.byref(r))
        self.energy_gs = d['energy']
        self.stab_gs = StabilizerGroup(_group_from_result(d, self.n), canonicalized=True)
        self.term_signs = d['term_signs']


class StateMachinePeriodic(StateMachine):
    """py/state.py:402-462: the same machine on the unrolled periodic chain (its right environments come from
    generate_Sright_periodic); the initial states (one per right environment at site 0) are set by the handle."""

    def __init__(self, hamiltonian, Sright_all):
        super().__init__(hamiltonian, Sright_all)
        self.l = self.n - self.k + 1                  # py/state.py:406 (not the period)
