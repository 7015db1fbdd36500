"""Per-launch metric table out of an .ncu-rep WITH UNITS (ncu -i rep --page raw --csv carries a units row; values are printed as
"<value> <unit>").  python tools/ncu_summary.py rep.ncu-rep [kernel-regex]"""
import csv
import re
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor_op_hmma.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum']


def main(rep, pattern='.'):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index('Kernel Name')
    for r in rows[2:]:
        if len(r) <= name_col or not re.search(pattern, r[name_col]):
            continue
        print(r[name_col][:110])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print('    %-78s %s %s' % (w, r[i], units[i]))
        for w in hdr:
            if 'tensor' in w and w not in WANT:
                i = hdr.index(w)
                print('    %-78s %s %s' % (w, r[i], units[i]))
        print()


if __name__ == '__main__':
    main(*sys.argv[1:])
