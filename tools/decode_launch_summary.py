"""Summarise the last decode step of an `ncu --metrics gpu__time_duration.sum --csv` launch list of tools/decode_profile.py
(the launches between the two spin_kernel markers).  usage: python tools/decode_launch_summary.py launches.csv"""
import collections
import csv
import re
import sys

with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith('==')]
rows = [(r['Kernel Name'], float(r['Metric Value'])) for r in csv.DictReader(lines)]
marks = [i for i, (n, _) in enumerate(rows) if 'spin_kernel' in n]
step = rows[marks[-2] + 1: marks[-1]]
tot = sum(v for _, v in step)
print(f'launches {len(step)}  total {tot / 1e3:.1f} us (cold-cache, serialised)')
agg = collections.defaultdict(lambda: [0, 0.0])
for n, v in step:
    n = re.sub(r'^void ', '', n)
    m = re.match(r'([\w:]+(<[^(]{0,60})?)', n)
    a = agg[(m.group(1) if m else n)[:90]]
    a[0] += 1
    a[1] += v
print('| share | time (us) | launches | kernel |\n|---:|---:|---:|---|')
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f'| {v / tot * 100:.1f}% | {v / 1e3:.1f} | {c} | `{k}` |')
