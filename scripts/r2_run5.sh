for sc in default; do
echo "== default (split from 5 shards, 2500 cliques)"; python scripts/shard_sweep.py 24 300 0 8 3 | tail -1
done
for sc in 1200 600 300 150; do echo "== SGS_SPLIT_CLIQUES=$sc"; SGS_SPLIT_CLIQUES=$sc python scripts/shard_sweep.py 24 300 0 8 3 | tail -1; done
echo "== NOSPLIT"; SGS_NOSPLIT=1 python scripts/shard_sweep.py 24 300 0 8 3 | tail -1
echo "== world 2/4, split 300"; SGS_SPLIT=1 SGS_SPLIT_CLIQUES=300 python scripts/shard_sweep.py 24 300 0 2 3 | tail -1; SGS_SPLIT=1 SGS_SPLIT_CLIQUES=300 python scripts/shard_sweep.py 24 300 0 4 3 | tail -1
echo "== world 2/4, nosplit"; python scripts/shard_sweep.py 24 300 0 2 3 | tail -1; python scripts/shard_sweep.py 24 300 0 4 3 | tail -1
echo "== world 1 split 300 (single GPU with rounds)"; SGS_SPLIT=1 SGS_SPLIT_CLIQUES=300 python scripts/run_sparse_once.py 24 300 0 4 | tail -1
cp build_ab/libsgs_trace.so stabilizer_gs_b200/libsgs.so
echo "== trace build, world 8"; SGS_TRACE=1 python scripts/shard_sweep.py 24 300 0 8 2 2>&1 | grep -A2 "replay shard 3" | tail -3
SGS_TRACE=1 SGS_SPLIT_CLIQUES=300 python scripts/shard_sweep.py 24 300 0 8 2 2>&1 | grep -A2 "replay shard 3" | tail -3
