# hand-off rounds only at the end of a launch (SGS_SPLIT_MARGIN roots per warp left): sweep on virtual shards
run() { echo "== $*"; env "${@:5}" timeout 300 python scripts/shard_sweep.py $1 $2 0 $3 $4 2>&1 | tail -1 | sed 's/.*| //'; }
for mg in 0 1 2 4 8; do
  run 24 300 8 6 SGS_SPLIT_MARGIN=$mg
done
for mg in 0 1 2 4 8; do
  run 24 300 4 6 SGS_SPLIT=1 SGS_SPLIT_MARGIN=$mg
done
run 24 300 4 6 SGS_NOSPLIT=1
for mg in 0 2 4; do
  run 24 300 2 6 SGS_SPLIT=1 SGS_SPLIT_MARGIN=$mg
done
run 24 300 2 6 SGS_NOSPLIT=1
for mg in 0 1 2 4; do
  run 40 500 8 3 SGS_SPLIT_MARGIN=$mg
done
run 40 500 8 3 SGS_NOSPLIT=1
for mg in 2 4; do
  run 40 500 2 3 SGS_SPLIT=1 SGS_SPLIT_MARGIN=$mg
done
run 40 500 2 3 SGS_NOSPLIT=1
run 30 400 8 3 SGS_SPLIT_MARGIN=0
run 30 400 8 3 SGS_SPLIT_MARGIN=2
run 30 400 8 3 SGS_NOSPLIT=1
