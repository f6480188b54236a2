"""Summarise an ncu launch list (gpu__time_duration.sum per launch) by kernel name."""
import collections
import csv
import sys

fn = sys.argv[1]
with open(fn) as f:
    lines = [l for l in f if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = row['Kernel Name'].split('(')[0][:64]
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    if u == 'ns':
        v /= 1e3
    elif u == 'ms':
        v *= 1e3
    elif u in ('s', 'second'):
        v *= 1e6
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print("%-66s %5s %12s %10s %7s" % ("kernel", "n", "total_us", "avg_us", "share"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-66s %5d %12.1f %10.1f %6.1f%%" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
