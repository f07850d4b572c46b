"""bench.py's roofline arithmetic on the committed ncu counters (no GPU): the utilisation it prints is a fraction, the
counters of a shard capture are scaled by the shard's share of the candidates, N GPUs divide by N peaks."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

PEAKS = dict(lop3_ops_per_s=18.53e12, popc_ops_per_s=4.64e12, sm_mhz=1956.0)


def _counters():
    return json.load(open(os.path.join(ROOT, 'profiles', 'r02_k_dfs_counters.json')))


def test_committed_counters_cover_both_benchmark_workloads():
    c = _counters()
    assert 'config4' in c and 'config5' in c
    assert c['config4']['instance_scale'] is None
    assert 0.0 < float(c['config5']['instance_scale']) < 1.0          # one virtual shard of the workload


def test_roofline_fraction_is_a_fraction_and_scales_with_gpus():
    c = _counters()
    hist = {3: 1000}
    # live kernel time = the capture's own time: the fraction must reproduce ncu's ALU-pipe utilisation (within the
    # tail of the launch, sm__cycles_active / elapsed) and stay below 1
    t4 = c['config4']['gpu_time_ms'] * 1e-3
    b1 = bench.roofline_block('config4', t4, 1, PEAKS, hist, 300, 24)
    assert 0.0 < b1['frac'] < 1.0 and 0.0 < b1['issue_slot_frac'] < 1.0 and b1['frac_lane_weighted'] < b1['frac']
    assert abs(b1['frac'] - c['config4']['alu_pct_of_peak_ncu'] / 100.0) < 0.05
    b8 = bench.roofline_block('config4', t4, 8, PEAKS, hist, 300, 24)
    assert abs(b8['frac'] * 8 - b1['frac']) < 1e-12                    # eight GPUs: eight peaks
    assert b1['algorithmic_over_peak'] > 0 and 'not a utilisation' in b1['algorithmic_note']


def test_shard_capture_is_scaled_by_its_share():
    c = _counters()['config5']
    share = float(c['instance_scale'])
    t = c['gpu_time_ms'] * 1e-3 / share                                 # the whole workload at the shard's rate
    b = bench.roofline_block('config5', t, 1, PEAKS, {7: 1000}, 1000, 40)
    assert abs(b['frac'] - c['alu_pct_of_peak_ncu'] / 100.0) < 0.05
    assert abs(b['traffic'] - c['dram_bytes'] / share) < 1.0
