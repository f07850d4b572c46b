# roots one level higher / lower (existing level A does the rest), with the heavy-first order
for d in 0 2 4; do for p in 0 1; do
  echo "== config 4 expand_depth=$d perm=$p"; if [ $p = 1 ]; then export SGS_PERM=1; else unset SGS_PERM; fi
  timeout 120 python scripts/run_sparse_once.py 24 300 0 4 $d | tail -1 | sed 's/.*slow=/slow=/'
done; done
for d in 0 2 4; do
  echo "== n=40 T=500 expand_depth=$d perm=1"; SGS_PERM=1 timeout 120 python scripts/run_sparse_once.py 40 500 0 3 $d | tail -1 | sed 's/.*slow=/slow=/'
done
echo "== n=40 T=500 perm=0"; unset SGS_PERM; timeout 120 python scripts/run_sparse_once.py 40 500 0 3 | tail -1 | sed 's/.*slow=/slow=/'
