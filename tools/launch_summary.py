"""Per-kernel totals of an ncu launch list (csv from `ncu --metrics gpu__time_duration.sum --csv`): python tools/launch_summary.py file.csv"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
tot = collections.defaultdict(float); cnt = collections.Counter(); mx = collections.defaultdict(float)
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'<.*', '', row['Kernel Name']).split('(')[0].replace('void ', '')
    v = float(row['Metric Value'].replace(',', ''))
    v = v / 1e3 if row['Metric Unit'] == 'ns' else (v * 1e3 if row['Metric Unit'] == 'ms' else v)
    tot[name] += v; cnt[name] += 1; mx[name] = max(mx[name], v)
all_ = sum(tot.values())
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"{k:36s} n={cnt[k]:6d} total {tot[k] / 1e3:10.2f} ms ({100 * tot[k] / all_:5.1f} %)  mean {tot[k] / cnt[k]:9.1f} us  max {mx[k]:9.1f} us")
print(f"{'all':36s} n={sum(cnt.values()):6d} total {all_ / 1e3:10.2f} ms")
