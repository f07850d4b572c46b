"""ctypes binding of libsgs.so (include/sgs.h).  No torch, no numpy requirement beyond arrays.

The library is built in-tree by `__graft_entry__.build()` (or `make -C stabilizer_gs_b200/csrc`).
Loading fails loudly if it is missing: there is no Python/CPU fallback for any search.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libsgs.so')

SGS_MAX_RANK = 64
SGS_NCCL_ID_BYTES = 128

ERRORS = {
    0: 'SGS_OK', -1: 'SGS_ERR_IO', -2: 'SGS_ERR_PARSE', -3: 'SGS_ERR_ARG', -4: 'SGS_ERR_NO_DEVICE',
    -5: 'SGS_ERR_CUDA', -6: 'SGS_ERR_LIMIT', -7: 'SGS_ERR_NCCL', -8: 'SGS_ERR_OVERFLOW', -9: 'SGS_ERR_DEAD',
}


class SgsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


class Opts(C.Structure):
    _fields_ = [('device', C.c_int), ('ngpus', C.c_int), ('rank', C.c_int), ('world', C.c_int),
                ('nccl_id', C.c_void_p), ('verbose', C.c_int), ('expand_depth', C.c_int), ('prune', C.c_int),
                ('reserved', C.c_int * 7)]


class Result(C.Structure):
    _fields_ = [('energy', C.c_double), ('n', C.c_int), ('m', C.c_int), ('W', C.c_int), ('T', C.c_int),
                ('gens', C.POINTER(C.c_uint64)), ('gen_signs', C.POINTER(C.c_int8)),
                ('term_signs', C.POINTER(C.c_int8)),
                ('candidates', C.c_uint64), ('rank_hist', C.c_uint64 * (SGS_MAX_RANK + 1)),
                ('sum_2m', C.c_double), ('slow_nodes', C.c_uint64),
                ('nsites', C.c_int), ('sright_counts', C.POINTER(C.c_int)),
                ('states_per_site', C.POINTER(C.c_int)), ('transitions', C.c_uint64),
                ('seconds_total', C.c_double), ('seconds_device', C.c_double),
                ('seconds_hot_kernel', C.c_double), ('kernel_launches', C.c_uint32),
                ('walk_paths', C.c_uint32 * 8)]


class Orbit(C.Structure):
    _fields_ = [('energy', C.c_double), ('begin', C.c_int), ('end', C.c_int), ('m', C.c_int),
                ('gens', C.POINTER(C.c_uint64)), ('signs', C.POINTER(C.c_int8))]


class PeriodicResult(C.Structure):
    _fields_ = [('l', C.c_int), ('k', C.c_int), ('n', C.c_int), ('W', C.c_int), ('nterms', C.c_int),
                ('nlog', C.c_int), ('sright_log', C.POINTER(C.c_int)),
                ('nfixed', C.c_int), ('fixed_sizes', C.POINTER(C.c_int)),
                ('norbits', C.c_int), ('orbits', C.POINTER(Orbit)), ('seconds_total', C.c_double)]


# every symbol include/sgs.h declares: name -> (restype, argtypes)
SYMBOLS = {
    'sgs_ham_load': (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    'sgs_ham_from_text': (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    'sgs_ham_from_arrays': (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_double),
                                      C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_void_p)]),
    'sgs_ham_free': (None, [C.c_void_p]),
    'sgs_ham_n': (C.c_int, [C.c_void_p]),
    'sgs_ham_k': (C.c_int, [C.c_void_p]),
    'sgs_ham_l': (C.c_int, [C.c_void_p]),
    'sgs_ham_nterms': (C.c_int, [C.c_void_p]),
    'sgs_ham_words': (C.c_int, [C.c_void_p]),
    'sgs_ham_terms': (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    'sgs_ham_order': (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    'sgs_opts_default': (None, [C.POINTER(Opts)]),
    'sgs_result_free': (None, [C.POINTER(Result)]),
    'sgs_sparse_search': (C.c_int, [C.c_void_p, C.POINTER(Opts), C.POINTER(Result)]),
    'sgs_sparse_plan_create': (C.c_int, [C.c_void_p, C.POINTER(Opts), C.POINTER(C.c_void_p)]),
    'sgs_sparse_plan_run': (C.c_int, [C.c_void_p, C.POINTER(Result)]),
    'sgs_sparse_plan_free': (None, [C.c_void_p]),
    'sgs_klocal_search': (C.c_int, [C.c_void_p, C.POINTER(Opts), C.POINTER(Result)]),
    'sgs_klocal_sright_log': (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.c_int]),
    'sgs_klocal_sright_group': (C.c_int, [C.c_void_p, C.c_int, C.c_int] + [C.POINTER(C.c_int)] * 3 + [C.POINTER(C.c_uint64), C.POINTER(C.c_int8)]),
    'sgs_klocal_state_valid_terms': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int)]),
    'sgs_klocal_get_states': (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    'sgs_klocal_states_count': (C.c_int, [C.c_void_p]),
    'sgs_klocal_states_free': (None, [C.c_void_p]),
    'sgs_klocal_states_select': (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]),
    'sgs_klocal_states_append': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    'sgs_klocal_state_info': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int] + [C.POINTER(C.c_int)] * 5),
    'sgs_klocal_state_rows': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_uint64)]),
    'sgs_klocal_state_key': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_int)]),
    'sgs_klocal_state_energies': (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double)]),
    'sgs_klocal_state_set_energies': (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double)]),
    'sgs_klocal_states_shift': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    'sgs_klocal_set_states': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    'sgs_klocal_state_history': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int8), C.c_int,
                                           C.POINTER(C.c_int)]),
    'sgs_klocal_create': (C.c_int, [C.c_void_p, C.POINTER(Opts), C.c_int, C.POINTER(C.c_void_p)]),
    'sgs_klocal_free': (None, [C.c_void_p]),
    'sgs_klocal_m': (C.c_int, [C.c_void_p]),
    'sgs_klocal_nstates': (C.c_int, [C.c_void_p]),
    'sgs_klocal_sright_count': (C.c_int, [C.c_void_p, C.c_int]),
    'sgs_klocal_evolve': (C.c_int, [C.c_void_p]),
    'sgs_klocal_ground_state': (C.c_int, [C.c_void_p, C.POINTER(Result)]),
    'sgs_klocal_periodic': (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.POINTER(Opts), C.POINTER(PeriodicResult)]),
    'sgs_periodic_result_free': (None, [C.POINTER(PeriodicResult)]),
    'sgs_group_canonicalize': (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_int8), C.POINTER(C.c_int)]),
    'sgs_group_energy': (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_int8), C.c_int,
                                   C.POINTER(C.c_uint64), C.POINTER(C.c_int8), C.POINTER(C.c_int8)]),
    'sgs_pauli_commute': (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint8)]),
    'sgs_nccl_unique_id': (C.c_int, [C.c_void_p]),
    'sgs_device_count': (C.c_int, []),
    'sgs_last_error': (C.c_char_p, []),
    'sgs_version': (C.c_char_p, []),
    'sgs_device_flush_l2': (C.c_int, [C.c_int]),
    'sgs_measure_int_peaks': (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}

_lib = None


def lib():
    """Load libsgs.so once.  Raises (never falls back) if the CUDA library is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(stabilizer_gs_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)          # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise SgsError(rc, (lib().sgs_last_error() or b'').decode())


def make_opts(device=0, ngpus=1, rank=0, world=1, nccl_id=None, verbose=0, expand_depth=0, prune=0):
    o = Opts()
    lib().sgs_opts_default(C.byref(o))
    o.device, o.ngpus, o.rank, o.world, o.verbose, o.expand_depth = device, ngpus, rank, world, verbose, expand_depth
    o.prune = int(prune)
    keep = None
    if nccl_id is not None:
        keep = C.create_string_buffer(bytes(nccl_id), SGS_NCCL_ID_BYTES)
        o.nccl_id = C.cast(keep, C.c_void_p)
    return o, keep


def row_to_letters(words, n, begin=0, end=None):
    end = n if end is None else end
    return ''.join('IZXY'[(words[(2 * j) >> 6] >> ((2 * j) & 63)) & 3] for j in range(begin, end))


def result_to_dict(r, n):
    W = r.W
    gens = []
    for i in range(r.m):
        words = [r.gens[i * W + k] for k in range(W)]
        gens.append(('+' if r.gen_signs[i] >= 0 else '-') + row_to_letters(words, n))
    d = dict(energy=r.energy, n=r.n, m=r.m, generators=gens,
             term_signs=[int(r.term_signs[t]) for t in range(r.T)],
             candidates=int(r.candidates), sum_2m=float(r.sum_2m), slow_nodes=int(r.slow_nodes),
             rank_hist={i: int(r.rank_hist[i]) for i in range(SGS_MAX_RANK + 1) if r.rank_hist[i]},
             seconds_total=r.seconds_total, seconds_device=r.seconds_device,
             seconds_hot_kernel=r.seconds_hot_kernel, kernel_launches=int(r.kernel_launches),
             transitions=int(r.transitions), walk_paths=[int(r.walk_paths[i]) for i in range(8)])
    if r.nsites:
        d['sright_counts'] = [int(r.sright_counts[i]) for i in range(r.nsites + 1)]
        d['states_per_site'] = [int(r.states_per_site[i]) for i in range(r.nsites + 1)]
    return d


class HamHandle:
    """Owns an sgs_ham*."""

    def __init__(self, ptr):
        self.ptr = ptr

    @classmethod
    def load(cls, path, periodic=False, cmax=0):
        p = C.c_void_p()
        check(lib().sgs_ham_load(os.fsencode(path), int(periodic), int(cmax), C.byref(p)))
        return cls(p)

    @classmethod
    def from_text(cls, text, periodic=False, cmax=0):
        p = C.c_void_p()
        check(lib().sgs_ham_from_text(text.encode(), int(periodic), int(cmax), C.byref(p)))
        return cls(p)

    @property
    def n(self): return lib().sgs_ham_n(self.ptr)
    @property
    def k(self): return lib().sgs_ham_k(self.ptr)
    @property
    def l(self): return lib().sgs_ham_l(self.ptr)
    @property
    def nterms(self): return lib().sgs_ham_nterms(self.ptr)
    @property
    def words(self): return lib().sgs_ham_words(self.ptr)

    def terms(self):
        T, W = self.nterms, self.words
        zx = (C.c_uint64 * max(T * W, 1))()
        w = (C.c_double * max(T, 1))()
        b = (C.c_int * max(T, 1))()
        e = (C.c_int * max(T, 1))()
        check(lib().sgs_ham_terms(self.ptr, zx, w, b, e))
        return ([[zx[t * W + k] for k in range(W)] for t in range(T)], list(w)[:T], list(b)[:T], list(e)[:T])

    def order(self):
        T = self.nterms
        o = (C.c_int * max(T, 1))()
        check(lib().sgs_ham_order(self.ptr, o))
        return list(o)[:T]

    def __del__(self):
        if getattr(self, 'ptr', None) and _lib is not None:
            _lib.sgs_ham_free(self.ptr)
            self.ptr = None


def sparse_search(ham, **kw):
    """FullState(paulis).evolve() on the GPU.  `ham` is a HamHandle.  Returns a dict."""
    o, keep = make_opts(**kw)
    r = Result()
    check(lib().sgs_sparse_search(ham.ptr, C.byref(o), C.byref(r)))
    try:
        return result_to_dict(r, ham.n)
    finally:
        lib().sgs_result_free(C.byref(r))


class SparsePlan:
    """Device-resident term tables; run() repeats the search without re-uploading."""

    def __init__(self, ham, **kw):
        o, self._keep = make_opts(**kw)
        self.ham = ham
        self.ptr = C.c_void_p()
        check(lib().sgs_sparse_plan_create(ham.ptr, C.byref(o), C.byref(self.ptr)))

    def run(self):
        r = Result()
        check(lib().sgs_sparse_plan_run(self.ptr, C.byref(r)))
        try:
            return result_to_dict(r, self.ham.n)
        finally:
            lib().sgs_result_free(C.byref(r))

    def close(self):
        if self.ptr:
            lib().sgs_sparse_plan_free(self.ptr)
            self.ptr = None

    def __del__(self):
        if _lib is not None:
            self.close()


def klocal_search(ham, **kw):
    o, keep = make_opts(**kw)
    r = Result()
    check(lib().sgs_klocal_search(ham.ptr, C.byref(o), C.byref(r)))
    try:
        return result_to_dict(r, ham.n)
    finally:
        lib().sgs_result_free(C.byref(r))


def klocal_periodic(path, cmax=10, c=3, **kw):
    o, keep = make_opts(**kw)
    r = PeriodicResult()
    check(lib().sgs_klocal_periodic(os.fsencode(path), cmax, c, C.byref(o), C.byref(r)))
    try:
        W = r.W
        orbits = []
        for i in range(r.norbits):
            ob = r.orbits[i]
            gens = []
            for g in range(ob.m):
                words = [ob.gens[g * W + k] for k in range(W)]
                gens.append(('+' if ob.signs[g] >= 0 else '-') + row_to_letters(words, r.n, ob.begin, ob.end))
            orbits.append(dict(energy=ob.energy, begin=ob.begin, end=ob.end, generators=gens))
        return dict(l=r.l, k=r.k, n=r.n, nterms=r.nterms,
                    sright_log=[[r.sright_log[2 * i], r.sright_log[2 * i + 1]] for i in range(r.nlog)],
                    fixed_point_sizes=[r.fixed_sizes[i] for i in range(r.nfixed)], results=orbits,
                    seconds_total=r.seconds_total)
    finally:
        lib().sgs_periodic_result_free(C.byref(r))


def nccl_unique_id():
    buf = C.create_string_buffer(SGS_NCCL_ID_BYTES)
    check(lib().sgs_nccl_unique_id(buf))
    return buf.raw


def flush_l2(device=0):
    check(lib().sgs_device_flush_l2(device))


def measure_int_peaks(device=0):
    a, b, c = C.c_double(), C.c_double(), C.c_double()
    check(lib().sgs_measure_int_peaks(device, C.byref(a), C.byref(b), C.byref(c)))
    return dict(lop3_ops_per_s=a.value, popc_ops_per_s=b.value, sm_mhz=c.value)
