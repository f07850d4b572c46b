#!/bin/bash
# A/B of the heavy-first root order at N GPUs (bench.py value + hot kernel time)
cd "$(dirname "$0")/.."
N=${1:-2}
for v in 0 1; do
  if [ $v = 1 ]; then export SGS_NOPERM=1; else unset SGS_NOPERM; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$v bench.py --gpus $N --steps 20 --warmup 3 2>gpurun_out/ab_$v.err | tail -1 > gpurun_out/ab_$v.json
  python -c "
import json; d=json.loads(open('gpurun_out/ab_$v.json').read())
print('noperm=$v N=$N value=%.4g ms/step=%.3f hot=%.3f e2e_ms=%.3f' % (d['value'], d['ms_per_step'], d['roofline']['hot_kernel_ms'], d['e2e']['ms_per_step']))"
done
