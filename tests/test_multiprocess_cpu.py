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
sys.path.insert(0, sys.argv[2]); sys.path.insert(0, os.path.join(sys.argv[2], 'tests'))
import torch
import torch.distributed as dist
import stabilizer_gs_b200 as S
import gen_hamiltonians as G
rank = int(sys.argv[1])
dist.init_process_group('gloo', init_method='tcp://127.0.0.1:29534', rank=rank, world_size=2)
# the courier of bench.py: rank 0 creates the library's NCCL id (128 bytes), torch.distributed ships it
if rank == 0:
    try:
        nid = S.nccl_unique_id()
    except S.SgsError:                       # no NCCL on this host: any 128 bytes exercise the courier
        nid = bytes(range(128))
else:
    nid = None
ids = [nid]
dist.broadcast_object_list(ids, src=0)
assert isinstance(ids[0], (bytes, bytearray)) and len(ids[0]) == 128
# every rank binds the library with its rank / world / id: without a CUDA device the search must fail the same way on
# every rank (SGS_ERR_NO_DEVICE, no CPU fallback, no hang in a collective)
H = S.HamHandle.from_text(G.sparse_text(6, 12, rank))
o, keep = S._capi.make_opts(device=0, rank=rank, world=2, nccl_id=ids[0])
assert o.rank == rank and o.world == 2
code = 0
if S._capi.lib().sgs_device_count() == 0:
    for call in (lambda: S.sparse_search(H, rank=rank, world=2, nccl_id=ids[0]), lambda: S.klocal_search(H, rank=rank, world=2, nccl_id=ids[0])):
        try:
            call()
            code = 1
        except S.SgsError as e:
            code = e.code
        assert code == -4, code
# max over ranks of a per-rank time, as bench.py takes it
t = torch.tensor([float(rank + 1)], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert t.item() == 2.0
dist.barrier()
print('ok', rank, code)
'''


def test_library_under_a_two_rank_launcher_gloo_world2(tmp_path):
    """two processes (gloo, 127.0.0.1) load libsgs, ship the library's NCCL id the way bench.py does, bind their rank and
    world, and — on a host without GPU — get SGS_ERR_NO_DEVICE from both searches on both ranks"""
    script = tmp_path / 'w.py'
    script.write_text(_WORKER)
    ps = [subprocess.Popen([sys.executable, str(script), str(r), ROOT], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
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
