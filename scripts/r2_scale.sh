#!/bin/bash
# bench.py at 1 and N GPUs (torchrun, NCCL): scripts/r2_scale.sh N [steps]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}; K=${2:-20}
nvidia-smi -L | wc -l
timeout 600 python bench.py --steps $K --warmup 3 2> gpurun_out/r2s_n1.err | tail -1 > gpurun_out/r2s_n1.json; echo "n1 rc=$?"
for n in $(echo 2 4 8); do
  [ $n -le $N ] || continue
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps $K --warmup 3 2> gpurun_out/r2s_n$n.err | tail -1 > gpurun_out/r2s_n$n.json; echo "n$n rc=$?"
  tail -3 gpurun_out/r2s_n$n.err
done
python - <<PY
import json
base=None; b5=None
for N in (1,2,4,8):
    try: d=json.loads(open(f"gpurun_out/r2s_n{N}.json").read().strip().splitlines()[-1])
    except Exception as e: print(N,"missing",e); continue
    if N==1: base=d["value"]; b5=d.get('config5',{}).get('seconds')
    c5=d.get('config5',{})
    print(f"N={N} value={d['value']:.4g} cand/s ms/step={d['ms_per_step']:.3f} hot={d['roofline']['hot_kernel_ms']:.3f} e2e_ms={d['e2e']['ms_per_step']:.3f} speedup={d['value']/base:.2f} eff={d['value']/base/N:.2f} | config5: {c5.get('seconds')} s cand={c5.get('candidates')} speedup={(b5/c5['seconds']) if c5.get('seconds') and b5 else None}")
PY
