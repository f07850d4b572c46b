# host-side phases of the one-shot call at 2 GPUs (SGS_TRACE lines of both ranks)
SGS_TRACE=1 SGS_BENCH_NO_CONFIG5=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 2> gpurun_out/e2e2.err | tail -1 | cut -c1-300
grep "sgs_sparse_search:" gpurun_out/e2e2.err | tail -8
grep -c "" gpurun_out/e2e2.err
