#!/bin/bash
# bench.py at 8 and 2 GPUs (torchrun, NCCL) on one 8-GPU box; the N=1 line comes from a 1-GPU call
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in ${SCALE_NS:-8 2 4}; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 2> gpurun_out/r2w_n$n.err | tail -1 > gpurun_out/r2w_n$n.json; echo "n$n rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2w_n$n.json").read().strip().splitlines()[-1])
c5=d.get('config5',{})
print(f"N=$n value={d['value']:.4g} cand/s ms/step={d['ms_per_step']:.3f} hot={d['roofline']['hot_kernel_ms']:.3f} e2e_ms={d['e2e']['ms_per_step']:.3f} | config5: {c5.get('seconds')} s cand={c5.get('candidates')} E={c5.get('energy')}")
PY
done
