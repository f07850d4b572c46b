import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import stabilizer_gs_b200 as S
import gen_hamiltonians as G
H = S.HamHandle.from_text(G.klocal_text(64, 4, 3, 0))
ref = S.klocal_search(H)
for env in ({'SGS_DP_CAPS': '8'}, {'SGS_DP_CAPS': '8', 'SGS_SR_CAP': '32'}, {'SGS_DP_CAPS': '8', 'SGS_SR_POOL_KB': '8'}, {'SGS_DP_CAPS': '8', 'SGS_SR_CAP': '32', 'SGS_SR_POOL_KB': '8'}):
    for k in ('SGS_DP_CAPS', 'SGS_SR_CAP', 'SGS_SR_POOL_KB'):
        os.environ.pop(k, None)
    os.environ.update(env)
    r = S.klocal_search(H)
    d = [(i, a, b) for i, (a, b) in enumerate(zip(ref['states_per_site'], r['states_per_site'])) if a != b]
    e = [(i, a, b) for i, (a, b) in enumerate(zip(ref['sright_counts'], r['sright_counts'])) if a != b]
    print(env, 'transitions', r['transitions'], 'sum sp', sum(r['states_per_site'][:-1]), 'diff states', d[:6], 'diff sright', e[:6], 'E', r['energy'])
