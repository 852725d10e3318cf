"""Summarise ncu outputs brought back in gpurun_out/ into the small CSVs committed under profiles/.

    python profiles/summarize.py launches gpurun_out/launches.csv profiles/rNN_x_launches_summary.csv
    python profiles/summarize.py full gpurun_out/prof.ncu-rep profiles/rNN_x_ncu_full.csv
"""
import csv
import subprocess
import sys
from collections import defaultdict

KEEP = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum',
        'smsp__pcsamp_warps_issue_stalled_long_scoreboard', 'smsp__pcsamp_warps_issue_stalled_short_scoreboard',
        'smsp__pcsamp_warps_issue_stalled_barrier', 'smsp__pcsamp_warps_issue_stalled_mio_throttle',
        'smsp__pcsamp_warps_issue_stalled_lg_throttle', 'smsp__pcsamp_warps_issue_stalled_wait',
        'smsp__pcsamp_warps_issue_stalled_not_selected', 'smsp__pcsamp_warps_issue_stalled_branch_resolving',
        'smsp__pcsamp_warps_issue_stalled_no_instructions', 'smsp__pcsamp_warps_issue_stalled_math_pipe_throttle',
        'smsp__pcsamp_warps_issue_stalled_membar', 'smsp__pcsamp_warps_issue_stalled_dispatch_stall',
        'smsp__pcsamp_warps_issue_stalled_selected', 'smsp__pcsamp_sample_buffer_full']


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(',', ''))
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(r[ui], 1.0)
        agg[r[ki]][0] += 1
        agg[r[ki]][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, 'w') as f:
        w = csv.writer(f)
        w.writerow(['kernel', 'launches', 'total_us', 'avg_us', 'share_pct'])
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, v[0], f'{v[1]:.1f}', f'{v[1] / v[0]:.1f}', f'{100 * v[1] / tot:.2f}'])
    print(open(dst).read())


def full(src, dst):
    raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, 'w') as f:
        w = csv.writer(f)
        w.writerow(['metric', 'unit'] + [f'launch{i}' for i in range(len(rows) - 2)])
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                w.writerow([k, units[i]] + [r[i] for r in rows[2:]])
    print(open(dst).read())


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2], sys.argv[3])
