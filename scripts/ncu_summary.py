#!/usr/bin/env python
"""One-screen summary of an ncu report (first profiled kernel): python scripts/ncu_summary.py rep [rep ...]"""
import csv
import io
import subprocess
import sys

KEYS = [
    ('gpu__time_duration.sum', 'duration'),
    ('launch__grid_size', 'grid'), ('launch__block_size', 'block'), ('launch__registers_per_thread', 'regs/thread'),
    ('launch__occupancy_limit_shared_mem', 'CTAs/SM (smem limit)'), ('launch__occupancy_limit_registers', 'CTAs/SM (reg limit)'),
    ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput %'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 throughput %'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
    ('smsp__inst_executed.sum', 'warp instructions'),
    ('smsp__thread_inst_executed_per_inst_executed.ratio', 'threads / instruction'),
    ('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'FP64 pipe %'),
    ('sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe (hmma) active %'),
    ('TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'tensor pipe active % (realtime)'),
    ('sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor memory path active %'),
    ('sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed', 'bf16 tensor ops % of peak'),
    ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall long_scoreboard / issue'),
    ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'stall short_scoreboard / issue'),
    ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'stall barrier / issue'),
    ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'stall wait / issue'),
    ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'stall math throttle / issue'),
]


def summarize(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], check=True, stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL).stdout.decode()
    rows = list(csv.reader(io.StringIO(raw)))
    h, u = rows[0], rows[1]
    out = []
    for v in rows[2:]:
        d = dict(zip(h, zip(u, v)))
        out.append('### {}\n\nkernel: `{}`\n'.format(path, d['Kernel Name'][1]))
        out.append('| metric | value |\n|---|---|')
        for key, label in KEYS:
            if key in d and d[key][1] != '':
                out.append('| {} | {} {} |'.format(label, d[key][1], d[key][0]))
        out.append('')
    return '\n'.join(out)


if __name__ == '__main__':
    for p in sys.argv[1:]:
        print(summarize(p))
