"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of the captured window)."""
import collections
import csv
import re
import sys


def main(path, header=''):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    rows = []
    for row in csv.DictReader(lines):
        if row.get('Metric Name') == 'gpu__time_duration.sum':
            try:
                v = float(row['Metric Value'].replace(',', ''))
            except ValueError:
                continue
            if v != v:   # a launch ncu could not time (n/a)
                continue
            rows.append((row['Kernel Name'], v, row['Grid Size'], row['Block Size']))
    if header:
        print(header)
    agg = collections.OrderedDict()
    for n, v, g, b in rows:
        k = re.sub(r'\(.*', '', n)
        agg.setdefault(k, [0, 0.0, g, b])
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-62s n=%3d total=%9.1f us avg=%8.1f us share=%5.1f%% grid=%s block=%s' % (k[:62], v[0], v[1] / 1e3, v[1] / v[0] / 1e3,
                                                                                        100 * v[1] / tot, v[2], v[3]))
    print('total %.1f us over %d launches' % (tot / 1e3, len(rows)))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else '')
