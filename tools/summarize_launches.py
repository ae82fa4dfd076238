"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_bench.txt"""
import collections
import csv
import re
import sys

rows = []
with open(sys.argv[1], newline='') as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(r['Metric Value'].replace(',', ''))
    unit = r['Metric Unit']
    v_us = v * {'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3, 'nsecond': 1e-3, 's': 1e6, 'second': 1e6}[unit]
    name = re.sub(r'\(.*$', '', r['Kernel Name']).replace('(anonymous namespace)::', '')
    name = r['Kernel Name'].replace('(anonymous namespace)::', '')
    name = re.sub(r'\((?:[^()]|\([^()]*\))*\)\s*$', '', name).replace('void ', '')
    rows.append((name, v_us))
tot = sum(v for _, v in rows)
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for n, v in rows:
    a = agg[n]
    a[0] += 1
    a[1] += v
    a[2] = max(a[2], v)
print(f'# {sys.argv[1]}: {len(rows)} launches, {tot/1e3:.3f} ms of kernel time (ncu: serialised, cold caches -- shares only)')
print(f'{"kernel":70s} {"launches":>8s} {"total ms":>10s} {"share":>7s} {"max us":>10s}')
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{n[:70]:70s} {a[0]:8d} {a[1]/1e3:10.3f} {100*a[1]/tot:6.2f}% {a[2]:10.1f}')
