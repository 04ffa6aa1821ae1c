"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per (kernel, grid): count, mean, share.
Usage: python profiles/agg_launches.py launches.csv [skip_first_n]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        hdr, start = r, i
        break
ki, vi, gi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size')
agg = collections.OrderedDict()
n = 0
for r in rows[start + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    n += 1
    if n <= skip:
        continue
    name = r[ki].split('(')[0].replace('void ', '').replace('<unnamed>::', '')
    agg.setdefault((name[:48], r[gi]), []).append(v)
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print('%-50s %-14s n=%4d mean=%9.1f us  min=%8.1f max=%8.1f  %5.1f%%' % (k[0], k[1], len(v), sum(v) / len(v) / 1e3, min(v) / 1e3, max(v) / 1e3, 100 * sum(v) / tot))
print('total %.1f us over %d launches' % (tot / 1e3, n - skip))
