"""Hot CUDA source lines of one kernel in an ncu report (needs -lineinfo and --import-source on):
   python tools/ncu_lines.py report.ncu-rep 'regex' [top]"""
import csv
import subprocess
import sys


def main(rep, pattern, top=30):
    import re
    base = re.split(r'[<(]', pattern)[0]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + base, '--print-source', 'cuda,sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fname, hdr, func0, cur = None, None, None, None
    data = []
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            fname = r[1].split('/')[-1]
            continue
        if r[0] == 'Function Name':
            cur = r[1]
            if func0 is None and re.search(pattern, cur):
                func0 = cur
                print(func0[:120])
            continue
        if r[0] == 'Line No':
            hdr = r
            continue
        if hdr and func0 is not None and cur == func0 and len(r) >= len(hdr) and r[0] not in ('', ):
            try:
                s = int(r[hdr.index('# Samples')])
                ie = int(r[hdr.index('Instructions Executed')])
            except Exception:
                continue
            if s or ie:
                data.append((s, ie, fname, r[0], r[1].strip()))
    tot = sum(d[0] for d in data)
    toti = sum(d[1] for d in data)
    print('samples', tot, 'warp-instructions', toti)
    for s, ie, f, ln, src in sorted(data, key=lambda d: -d[0])[:top]:
        print('%5.1f%% smp %5.1f%% ins  %s:%s  %s' % (100.0 * s / max(tot, 1), 100.0 * ie / max(toti, 1), f, ln, src[:100]))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30)
