"""ctypes access to the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use this;
the product (stabilizer_gs_b200/) never does.
"""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, 'oracle')
# SGS_ORACLE_OW=12 selects the wide build (rows of 12 words, chains of up to 382 sites); the table sizes of the
# energy tables limit it to the k-local DP in practice
OW = int(os.environ.get('SGS_ORACLE_OW', '4'))
assert OW in (4, 12)
LIB = os.path.join(ORACLE_DIR, 'liboracle.so' if OW == 4 else 'liboracle_w12.so')
OMAXQ = OW * 32
OMAXM = OMAXQ


class ORow(C.Structure):
    _fields_ = [('b', C.c_uint64 * OW), ('q', C.c_int)]


class OMask(C.Structure):
    _fields_ = [('w', C.c_uint64 * ((OMAXM + 63) // 64))]


class OStab(C.Structure):
    _fields_ = [('m', C.c_int), ('begin', C.c_int), ('end', C.c_int), ('g', ORow * OMAXM)]


class OHam(C.Structure):
    _fields_ = [('n', C.c_int), ('k', C.c_int), ('l', C.c_int), ('T', C.c_int),
                ('P', C.POINTER(ORow)), ('w', C.POINTER(C.c_double)),
                ('begin', C.POINTER(C.c_int)), ('end', C.POINTER(C.c_int)),
                ('order', C.POINTER(C.c_int)), ('bucket_off', C.POINTER(C.c_int))]


class ORes(C.Structure):
    _fields_ = [('energy', C.c_double), ('n', C.c_int), ('m', C.c_int), ('gens', ORow * OMAXM),
                ('term_signs', C.POINTER(C.c_int8)), ('T', C.c_int),
                ('candidates', C.c_uint64), ('rank_hist', C.c_uint64 * (OMAXM + 1)), ('sum_2m', C.c_double),
                ('nsites', C.c_int), ('sright_counts', C.POINTER(C.c_int)),
                ('states_per_site', C.POINTER(C.c_int)), ('transitions', C.c_uint64), ('seconds', C.c_double)]


class OOrbit(C.Structure):
    _fields_ = [('energy', C.c_double), ('begin', C.c_int), ('end', C.c_int), ('m', C.c_int), ('gens', ORow * OMAXM)]


class OPeriodic(C.Structure):
    _fields_ = [('l', C.c_int), ('k', C.c_int), ('n', C.c_int), ('nterms', C.c_int),
                ('nlog', C.c_int), ('sright_log', C.POINTER(C.c_int)),
                ('nfixed', C.c_int), ('fixed_sizes', C.POINTER(C.c_int)),
                ('norbits', C.c_int), ('orbits', C.POINTER(OOrbit))]


_lib = None


def build():
    subprocess.run(['make', '-C', ORACLE_DIR, '-s'], check=True, stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        L.oham_load.restype = C.POINTER(OHam)
        L.oham_load.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_int]
        L.oham_from_text.restype = C.POINTER(OHam)
        L.oham_from_text.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_int]
        L.oham_free.argtypes = [C.POINTER(OHam)]
        L.oracle_sparse.restype = C.POINTER(ORes)
        L.oracle_sparse.argtypes = [C.POINTER(OHam), C.c_int]
        L.oracle_sparse_sample.restype = C.POINTER(ORes)
        L.oracle_sparse_sample.argtypes = [C.POINTER(OHam), C.c_int, C.c_uint64, C.c_double]
        L.oracle_sparse_stride.restype = C.POINTER(ORes)
        L.oracle_sparse_stride.argtypes = [C.POINTER(OHam), C.c_int, C.c_int, C.c_int]
        L.oracle_klocal.restype = C.POINTER(ORes)
        L.oracle_klocal.argtypes = [C.POINTER(OHam), C.c_int]
        L.ores_free.argtypes = [C.POINTER(ORes)]
        L.oracle_periodic.restype = C.POINTER(OPeriodic)
        L.oracle_periodic.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_int]
        L.operiodic_free.argtypes = [C.POINTER(OPeriodic)]
        L.orow_from_string.argtypes = [C.POINTER(ORow), C.c_char_p]
        L.orow_to_string.argtypes = [C.POINTER(ORow), C.c_int, C.c_int, C.c_char_p]
        L.orow_commute.argtypes = [C.POINTER(ORow), C.POINTER(ORow)]
        L.orow_dot.argtypes = [C.POINTER(ORow), C.POINTER(ORow), C.POINTER(ORow)]
        L.orow_sign.argtypes = [C.POINTER(ORow)]
        L.ostab_init.argtypes = [C.POINTER(OStab), C.c_int, C.c_int]
        L.ostab_canonicalize.argtypes = [C.POINTER(OStab), C.POINTER(OMask)]
        L.ostab_add.argtypes = [C.POINTER(OStab), C.POINTER(ORow)]
        L.ostab_decompose.argtypes = [C.POINTER(OStab), C.POINTER(ORow), C.POINTER(OMask), C.POINTER(C.c_int)]
        L.ostab_commute.argtypes = [C.POINTER(OStab), C.POINTER(ORow)]
        L.ostab_project.argtypes = [C.POINTER(OStab), C.c_int, C.c_int, C.POINTER(OStab)]
        ip = C.POINTER(C.c_int)
        L.olin_inv_list.argtypes = [C.c_int, ip]
        L.olin_get_order.argtypes = [ip, C.c_int, C.c_int, ip]
        L.olin_gauss.argtypes = [ip, ip, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.olin_solve.argtypes = [ip, ip, C.c_int, C.c_int, C.c_int, C.c_int, ip, ip]
        L.olin_kernel.argtypes = [ip, C.c_int, C.c_int, C.c_int, ip]
        L.olin_simplify_coset.argtypes = [ip, C.c_int, C.c_int, ip, C.c_int, ip]
        _lib = L
    return _lib


def row_str(row, begin, end):
    buf = C.create_string_buffer(OMAXQ + 4)
    lib().orow_to_string(C.byref(row), begin, end, buf)
    return buf.value.decode()


def row_from(s):
    r = ORow()
    n = lib().orow_from_string(C.byref(r), s.encode())
    assert n >= 0, s
    return r


def _load(text=None, path=None, periodic=False, cmax=0):
    err = C.create_string_buffer(256)
    if text is not None:
        H = lib().oham_from_text(text.encode(), int(periodic), cmax, err, 256)
    else:
        H = lib().oham_load(os.fsencode(path), int(periodic), cmax, err, 256)
    if not H:
        raise ValueError(err.value.decode())
    return H


def _res(H, r, klocal):
    rr = r.contents
    n = rr.n
    d = dict(energy=rr.energy, n=n, m=rr.m,
             generators=[row_str(rr.gens[i], 0, n) for i in range(rr.m)],
             term_signs=[int(rr.term_signs[t]) for t in range(rr.T)], seconds=rr.seconds)
    if klocal:
        d.update(sright_counts=[rr.sright_counts[i] for i in range(n + 1)],
                 states_per_site=[rr.states_per_site[i] for i in range(n + 1)], transitions=int(rr.transitions))
    else:
        d.update(candidates=int(rr.candidates), sum_2m=rr.sum_2m,
                 rank_hist={i: int(rr.rank_hist[i]) for i in range(OMAXM + 1) if rr.rank_hist[i]})
    return d


def sparse(text=None, path=None, threads=1):
    H = _load(text, path)
    r = lib().oracle_sparse(H, threads)
    try:
        return _res(H, r, False)
    finally:
        lib().ores_free(r)
        lib().oham_free(H)


def sparse_sample(text=None, path=None, threads=1, max_leaves=0, seconds=0.0):
    """bounded sample of the sparse walk: returns candidates visited and seconds"""
    H = _load(text, path)
    r = lib().oracle_sparse_sample(H, threads, max_leaves, seconds)
    try:
        return _res(H, r, False)
    finally:
        lib().ores_free(r)
        lib().oham_free(H)


def sparse_stride(text=None, path=None, threads=1, stride=16, offset=0):
    """uniform sample: complete walk of every stride-th depth-2 subtree; returns candidates visited and seconds"""
    H = _load(text, path)
    r = lib().oracle_sparse_stride(H, threads, stride, offset)
    try:
        return _res(H, r, False)
    finally:
        lib().ores_free(r)
        lib().oham_free(H)


def klocal(text=None, path=None, threads=1):
    H = _load(text, path)
    r = lib().oracle_klocal(H, threads)
    try:
        return _res(H, r, True)
    finally:
        lib().ores_free(r)
        lib().oham_free(H)


def periodic(path, cmax=10, c=3):
    err = C.create_string_buffer(256)
    r = lib().oracle_periodic(os.fsencode(path), cmax, c, err, 256)
    if not r:
        raise ValueError(err.value.decode())
    rr = r.contents
    try:
        out = []
        for i in range(rr.norbits):
            o = rr.orbits[i]
            out.append(dict(energy=o.energy, begin=o.begin, end=o.end,
                            generators=[row_str(o.gens[g], o.begin, o.end) for g in range(o.m)]))
        return dict(l=rr.l, k=rr.k, n=rr.n, nterms=rr.nterms,
                    sright_log=[[rr.sright_log[2 * i], rr.sright_log[2 * i + 1]] for i in range(rr.nlog)],
                    fixed_point_sizes=[rr.fixed_sizes[i] for i in range(rr.nfixed)], results=out)
    finally:
        lib().operiodic_free(r)
