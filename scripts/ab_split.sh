#!/bin/bash
# A/B of heavy-first order / hand-off rounds at N GPUs: scripts/ab_split.sh N
cd "$(dirname "$0")/.."
N=${1:-2}
for v in none nosplit noperm_nosplit; do
  unset SGS_NOSPLIT SGS_NOPERM
  [ $v = nosplit ] && export SGS_NOSPLIT=1
  [ $v = noperm_nosplit ] && export SGS_NOSPLIT=1 SGS_NOPERM=1
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29617 bench.py --gpus $N --steps 20 --warmup 3 2>gpurun_out/ab_$v.err | tail -1 > gpurun_out/ab_$v.json
  python -c "
import json; d=json.loads(open('gpurun_out/ab_$v.json').read())
print('$v N=$N value=%.4g ms/step=%.3f hot=%.3f e2e_ms=%.3f' % (d['value'], d['ms_per_step'], d['roofline']['hot_kernel_ms'], d['e2e']['ms_per_step']))"
done
