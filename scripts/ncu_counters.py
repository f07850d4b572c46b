#!/usr/bin/env python
"""Summarise `ncu --set full` captures of k_dfs for bench.py's roofline block.

    python scripts/ncu_counters.py config4=gpurun_out/prof_c4.ncu-rep config5=gpurun_out/prof_w2.ncu-rep ...

For each capture: `ncu -i REP --page raw --csv`, first k_dfs launch → profiles/r02_k_dfs_<label>_ncu_full_summary.csv
(metric,unit,value for the metrics the judge and DESIGN.md quote) and one entry of profiles/r02_k_dfs_counters.json
(per-launch counters bench.py turns into the executed-instruction roofline).  A label like `config5@0.25` records that
the capture ran a 1/4-size instance of that workload (counters are then used for ratios only, not scaled).
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed.sum', 'smsp__inst_executed_pipe_alu.sum', 'smsp__inst_executed_pipe_xu.sum',
        'smsp__inst_executed_pipe_lsu.sum', 'smsp__inst_executed_pipe_fp64.sum', 'smsp__inst_executed_pipe_fma.sum',
        'smsp__inst_executed_pipe_fmaheavy.sum', 'smsp__inst_executed_pipe_uniform.sum', 'smsp__inst_executed_pipe_cbu.sum',
        'smsp__inst_executed_pipe_adu.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.avg.per_cycle_active',
        'sm__cycles_elapsed.max', 'sm__cycles_active.avg', 'smsp__cycles_active.avg', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.per_cycle_active', 'smsp__thread_inst_executed_pred_on_per_inst_executed.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_cbu.sum.pct_of_peak_sustained_active', 'l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct',
        'l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_local_op_st.sum', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warp_latency_issue_stalled_wait.ratio',
        'smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio', 'smsp__average_warp_latency_issue_stalled_no_instruction.ratio',
        'smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio', 'smsp__average_warp_latency_issue_stalled_not_selected.ratio',
        'smsp__average_warp_latency_issue_stalled_branch_resolving.ratio', 'smsp__average_warp_latency_issue_stalled_barrier.ratio',
        'smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio', 'smsp__average_warp_latency_issue_stalled_lg_throttle.ratio',
        'smsp__average_warp_latency_issue_stalled_mio_throttle.ratio',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum', 'smsp__inst_executed_op_shared_ld.sum',
        'smsp__inst_executed_op_shared_st.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'local_load_spilling_requests', 'derived__smsp__inst_executed_op_branch_pct']


def raw_rows(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def main():
    counters_path = os.path.join(ROOT, 'profiles', 'r02_k_dfs_counters.json')
    try:
        counters = json.load(open(counters_path))
    except (OSError, ValueError):
        counters = {}
    for arg in sys.argv[1:]:
        label, rep = arg.split('=', 1)
        scale = None
        if '@' in label:
            label, scale = label.split('@')
        hdr, units, rows = raw_rows(rep)
        kcol = hdr.index('Kernel Name')
        row = next(r for r in rows if 'k_dfs' in r[kcol])
        val = {h: (row[i], units[i]) for i, h in enumerate(hdr)}

        def num(name):
            if name not in val:
                return float('nan')
            v, u = val[name]
            x = float(v.replace(',', ''))
            mult = {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'byte': 1.0, 'us': 1e-3, 'ms': 1.0, 'ns': 1e-6, 's': 1e3,
                    'msecond': 1.0, 'usecond': 1e-3, 'nsecond': 1e-6, 'second': 1e3}.get(u, 1.0)
            return x * mult

        with open(os.path.join(ROOT, 'profiles', f'r02_k_dfs_{label}_ncu_full_summary.csv'), 'w', newline='') as f:
            w = csv.writer(f)
            w.writerow(['metric', 'unit', 'value'])
            w.writerow(['Kernel Name', '', row[kcol]])
            for k in KEEP:
                if k in val:
                    w.writerow([k, val[k][1], val[k][0]])
        counters[label] = dict(
            kernel=row[kcol].split('(')[0], capture=os.path.basename(rep), instance_scale=scale,
            gpu_time_ms=num('gpu__time_duration.sum'),
            inst_executed=num('smsp__inst_executed.sum'),
            thread_inst_executed=num('smsp__inst_executed.sum') * num('smsp__thread_inst_executed_per_inst_executed.ratio'),
            # ALU-pipe warp instructions of the launch: ncu reports the pipe as % of its peak (64 lanes = 2 warp
            # instructions per cycle and SM) over the active cycles
            sm_cycles_active_avg=num('sm__cycles_active.avg'), sm_count=148,
            inst_executed_pipe_alu=num('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active') / 100.0 * 2.0
            * num('sm__cycles_active.avg') * 148,
            xu_pct_of_peak_ncu=num('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'),
            lsu_pct_of_peak_ncu=num('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'),
            warps_active_per_sm=num('sm__warps_active.avg.per_cycle_active'),
            alu_pct_of_peak_ncu=num('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'),
            issue_active_pct_ncu=num('smsp__issue_active.avg.pct_of_peak_sustained_active'),
            dram_bytes=num('dram__bytes_read.sum') + num('dram__bytes_write.sum'),
            registers=num('launch__registers_per_thread'))
        counters[label] = {k: (None if isinstance(v, float) and v != v else v) for k, v in counters[label].items()}
        print(label, json.dumps(counters[label]))
    with open(counters_path, 'w') as f:
        json.dump(counters, f, indent=1)


if __name__ == '__main__':
    main()
