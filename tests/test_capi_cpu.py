"""CPU-side checks of the C ABI: the library loads, exports every symbol include/sgs.h declares, parses
Hamiltonians like the reference, and refuses (loudly) to search without a CUDA device."""
import os
import re

import pytest

import gen_hamiltonians as G
from conftest import GOLDEN, ROOT


def S():
    import stabilizer_gs_b200 as S_
    return S_


def test_header_symbols_are_all_exported_and_bound():
    hdr = open(os.path.join(ROOT, 'include', 'sgs.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(sgs_[a-z0-9_]+)\s*\(', hdr))
    C = S()._capi
    L = C.lib()
    assert declared == set(C.SYMBOLS), declared ^ set(C.SYMBOLS)
    for name in declared:
        assert hasattr(L, name)


def test_parser_matches_reference_layout():
    H = S().HamHandle.load(os.path.join(ROOT, 'examples', 'hamiltonian_finite.txt'))
    assert (H.n, H.k, H.nterms, H.words) == (4, 3, 9, 1)
    rows, w, b, e = H.terms()
    assert w[0] == -0.7463647462 and (b[8], e[8]) == (1, 4)
    # "YXY 1 2 3": Y = z|x = 3 at sites 1 and 3, X = 2 at site 2
    assert rows[8][0] == (3 << 2) | (2 << 4) | (3 << 6)
    # paulis_list order = by last site, then file order (cpp/state.cpp:88-96)
    assert H.order() == [0, 1, 4, 2, 5, 7, 3, 6, 8]


def test_periodic_parser_replicates_terms():
    H = S().HamHandle.load(os.path.join(ROOT, 'examples', 'hamiltonian_periodic.txt'), periodic=True, cmax=10)
    assert (H.l, H.k, H.n, H.nterms) == (1, 3, 17, 31)


def test_parse_errors():
    Sg = S()
    for bad in ("", "x y\n", "2 2\n 1.0 QZ 0 1\n", "2 2\n 1.0 ZZ 0\n", "2 2\n 1.0 ZZ 0 5\n", "200 40\n 1.0 ZZ 0 39\n"):
        with pytest.raises(Sg.SgsError):
            Sg.HamHandle.from_text(bad)
    # n > 127 parses (the k-local DP runs at any n); rows come back at full width
    H = Sg.HamHandle.from_text("200 2\n 1.0 Z 0\n -0.5 XY 198 199\n")
    assert (H.n, H.words, H.nterms) == (200, 7, 2)
    rows, w, b, e = H.terms()
    assert rows[0][0] == 1 and rows[1][6] == (2 << (2 * 198 - 384)) | (3 << (2 * 199 - 384)) and (b[1], e[1]) == (198, 200)
    # the Python reference stops at the first blank line (py/state.py:41-43)
    H = Sg.HamHandle.from_text("2 2\n 1.0 ZZ 0 1\n\n 1.0 XX 0 1\n")
    assert H.nterms == 1


def test_no_cpu_fallback():
    import ctypes
    Sg = S()
    if Sg._capi.lib().sgs_device_count() > 0:
        pytest.skip('a CUDA device is present')
    H = Sg.HamHandle.from_text(G.sparse_text(4, 6, 0))
    with pytest.raises(Sg.SgsError) as ei:
        Sg.sparse_search(H)
    assert ei.value.code == -4      # SGS_ERR_NO_DEVICE


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, 'stabilizer_gs_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cpp', '.cu', '.cuh', '.hpp', '.h')) or f == 'Makefile':
                txt = open(os.path.join(dirpath, f), errors='ignore').read()
                assert 'oracle/' not in txt.replace('not the oracle', '') and 'liboracle' not in txt and 'sgs_oracle' not in txt, f
