"""One-shot sgs_sparse_search calls on config 4 with SGS_TRACE phase timing (host side of the e2e number)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
os.environ['SGS_TRACE'] = '1'
import stabilizer_gs_b200 as S
import gen_hamiltonians as G
H = S.HamHandle.from_text(G.sparse_text(24, 300, 0))
for i in range(4):
    t = time.perf_counter()
    r = S.sparse_search(H)
    print('one-shot wall %.3f ms, device %.3f ms' % ((time.perf_counter() - t) * 1e3, r['seconds_device'] * 1e3), flush=True)
