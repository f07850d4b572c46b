cat > /tmp/zz.py <<'PY'
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import stabilizer_gs_b200 as S
n = int(sys.argv[1]); kw = eval(sys.argv[2])
text = f"{n} 2\n" + ''.join(f" {-1.0 - 0.01 * i:.4f} ZZ {i} {i + 1}\n" for i in range(n - 1))
H = S.HamHandle.from_text(text)
try:
    r = S.sparse_search(H, **kw); print('result', r['energy'], r['m'], r['candidates'])
except S.SgsError as e:
    print('code', e.code, e)
PY
for args in "70 dict(prune=1)" "70 dict(prune=1,expand_depth=1)" "70 dict()" "30 dict(prune=1)"; do
  echo "=== $args"; SGS_TRACE=1 timeout 40 python /tmp/zz.py $args 2>&1 | tail -12
done
