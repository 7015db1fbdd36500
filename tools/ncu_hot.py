"""Per-kernel hot spots from an ncu report: ncu -i rep --page source --csv | python tools/ncu_hot.py [top]
Groups SASS by the CUDA source line (needs -lineinfo / --import-source on) when available, else prints hot SASS."""
import csv
import subprocess
import sys


def kernels(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
            'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
            'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
            'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active']
    idx = [hdr.index(w) if w in hdr else None for w in want]
    for r in rows[2:]:
        print(' | '.join('%s=%s' % (w.split('.')[0].replace('smsp__average_warps_issue_stalled_', 'stall_')[:28], r[i][:44]) for w, i in zip(want, idx) if i is not None))


def hot(rep, pattern, top=25):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + pattern, '--print-source', 'cuda,sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    data = []
    for r in rows:
        if r and r[0] == 'Kernel Name':
            if data:
                break
            print(r[1][:100])
            continue
        if r and (r[0] == 'Address' or r[0] == '#' or 'Source' in r):
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        try:
            data.append((int(r[hdr.index('# Samples')]), int(r[hdr.index('Instructions Executed')]), r[hdr.index('Source')].strip()))
        except Exception:
            pass
    tot = sum(d[0] for d in data)
    print('instructions', len(data), 'samples', tot, 'warp-inst', sum(d[1] for d in data))
    for i, (s, ie, src) in sorted(sorted(enumerate(data), key=lambda t: -t[1][0])[:top]):
        print('%5d %6d %5.1f%% %9d  %s' % (i, s, 100.0 * s / max(tot, 1), ie, src[:110]))


if __name__ == '__main__':
    if len(sys.argv) == 2:
        kernels(sys.argv[1])
    else:
        hot(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
