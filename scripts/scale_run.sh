#!/bin/bash
# 1/2/4/8-GPU scaling of bench.py + BASELINE config 5 (n=40, T=1000) on all GPUs of the box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
python bench.py --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/scale_n1.json
for N in 2 4 8; do
  [ $N -le $NG ] || continue
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 2>gpurun_out/scale_n$N.err | tail -1 > gpurun_out/scale_n$N.json
done
python - <<PY
import json,glob
base=None
for N in (1,2,4,8):
    try: d=json.loads(open(f"gpurun_out/scale_n{N}.json").read().strip().splitlines()[-1])
    except Exception as e: print(N,"missing",e); continue
    if N==1: base=d["value"]
    print(f"N={N} value={d['value']:.4g} cand/s ms/step={d['ms_per_step']:.3f} hot={d['roofline']['hot_kernel_ms']:.3f} e2e_ms={d['e2e']['ms_per_step']:.3f} speedup={d['value']/base:.2f}")
PY
python -c "
import sys; sys.path.insert(0,'tests'); import gen_hamiltonians as G; G.gen_sparse(40,1000,0,'/tmp/c5.txt')"
./stabilizer_gs_b200/bin/sparse --gpus $NG --stats /tmp/c5.txt > gpurun_out/config5_${NG}gpu.txt 2> gpurun_out/config5_${NG}gpu.stats
tail -4 gpurun_out/config5_${NG}gpu.txt | cut -c1-200; cat gpurun_out/config5_${NG}gpu.stats
