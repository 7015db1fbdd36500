"""Kernel sequence of the LAST whole step in an ncu launch list produced with tools/profile_step.py."""
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = [(re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '').replace('mgb::', ''), float(r['Metric Value']) / 1e3, r['Grid Size'], r['Block Size'])
        for r in csv.DictReader(lines) if r.get('Metric Name') == 'gpu__time_duration.sum']
idx = [i for i, r in enumerate(rows) if r[0] == 'k_prep_params']
a = idx[-1]
tot = sum(r[1] for r in rows[a:])
if len(sys.argv) > 2:
    print(sys.argv[2])
for n, v, g, b in rows[a:]:
    print('%-34s %8.1f us %5.1f%%  grid=%s block=%s' % (n[:34], v, 100 * v / tot, g, b))
print('step total %.1f us over %d launches' % (tot, len(rows) - a))
