for d in 0 2 3; do echo "depth $d: $(python scripts/run_sparse_once.py 24 300 0 5 $d | tail -1 | sed 's/.*slow=/slow=/')"; done
for d in 0 2 3; do echo "w2 depth $d: $(python scripts/run_sparse_once.py 40 500 0 3 $d | tail -1 | sed 's/.*slow=/slow=/')"; done
