#!/bin/bash
cd "$(dirname "$0")/.."
for t in 0 1 2 3; do
  for w in 1 8; do
    SGS_TUNE=$t python scripts/run_shard_once.py 24 300 0 $w 4 > /tmp/o.txt 2>&1
    echo "tune=$t $(sed -n 1p /tmp/o.txt)"
  done
done
