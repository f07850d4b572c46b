"""world_size-2 CPU checks of the multi-process plumbing (gloo, 127.0.0.1): the reference arm of bench.py under
torchrun (rank 0 prints one JSON line, rank 1 exits 0 without work) and the courier that ships the NCCL id."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_under_torchrun_world2():
    env = dict(os.environ, OMP_NUM_THREADS='2')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29533', os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '0']
    # keep the CPU sample short for the test
    env['SGS_BENCH_SAMPLE_S'] = '1.0'
    p = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['n_gpus'] == 2 and d['value'] > 0
    assert d['cpu_baseline']['kind'] == 'port' and d['e2e']['h2d_bytes_per_step'] == 0


_WORKER = r'''
import os, sys
import torch.distributed as dist
dist.init_process_group('gloo', init_method='tcp://127.0.0.1:29534', rank=int(sys.argv[1]), world_size=2)
ids = [bytes(range(128)) if dist.get_rank() == 0 else None]
dist.broadcast_object_list(ids, src=0)
assert ids[0] == bytes(range(128))
import torch
t = torch.tensor([float(dist.get_rank() + 1)], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert t.item() == 2.0
dist.barrier()
print('ok', dist.get_rank())
'''


def test_id_courier_and_max_over_ranks_gloo_world2(tmp_path):
    script = tmp_path / 'w.py'
    script.write_text(_WORKER)
    ps = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300) for p in ps]
    for p, (o, e) in zip(ps, outs):
        assert p.returncode == 0, e[-2000:]
        assert 'ok' in o


def test_clock_sampler_without_a_gpu_reports_no_samples():
    """bench.py samples the SM clock through NVML (thread) or nvidia-smi; with neither it must not raise and must
    say that it saw nothing (the bench line then carries clocks with samples = 0)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('bench_mod', os.path.join(ROOT, 'bench.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    s = mod.ClockSampler(0)
    s.start()
    out = s.stop()
    assert set(out) >= {'sm_mhz', 'sm_max_mhz', 'reasons', 'samples'}
    assert out['samples'] == 0 or out['sm_mhz'] is not None
