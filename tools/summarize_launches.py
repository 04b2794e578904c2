#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, share).
usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.summary.txt"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').replace('<unnamed>::', '')
        v = float(row['Metric Value'].replace(',', ''))
        v *= {'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 's': 1e6}.get(row['Metric Unit'], 1.0)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print('# %s' % path)
    print('%-40s %6s %12s %7s %11s' % ('kernel', 'n', 'total_us', 'share', 'avg_us'))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-40s %6d %12.1f %6.1f%% %11.1f' % (k, v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))


if __name__ == '__main__':
    main(sys.argv[1])
