"""DRAM traffic per launch of one kernel from an `ncu --page raw --csv` export: (dram__bytes_read.sum + dram__bytes_write.sum)
averaged over the captured launches whose name matches the regex; optionally merged into profiles/traffic.json.

    python tools/ncu_traffic.py raw.csv 'k_atom_bwd' [workload-name profiles/traffic.json]"""
import csv
import json
import re
import sys

UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}


def main(path, pattern, workload=None, out=None):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {name: hdr.index(name) for name in ('dram__bytes_read.sum', 'dram__bytes_write.sum')}
    name_col = hdr.index('Kernel Name')
    vals = []
    for r in rows[2:]:
        if len(r) <= max(col.values()) or not re.search(pattern, r[name_col]):
            continue
        tot = 0.0
        for c in col.values():
            tot += float(r[c]) * UNIT[units[c]]
        vals.append(tot)
    if not vals:
        print('no launch matches', pattern)
        return
    avg = sum(vals) / len(vals)
    print('%s: %d launches, DRAM read+write per launch: avg %.0f bytes (min %.0f, max %.0f)' % (pattern, len(vals), avg, min(vals), max(vals)))
    if workload and out:
        try:
            data = json.load(open(out))
        except Exception:
            data = {}
        data.setdefault(pattern, {})[workload] = int(avg)
        json.dump(data, open(out, 'w'), indent=1)


if __name__ == '__main__':
    main(*sys.argv[1:])
