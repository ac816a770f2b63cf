"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one training step, time per kernel.
usage: python tools/launch_summary.py gpurun_out/launches_eager.csv [step_index]"""
import collections
import csv
import re
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith('==')]
for row in csv.DictReader(lines):
    rows.append((row['Kernel Name'], float(row['Metric Value'])))
marks = [i for i, (n, _) in enumerate(rows) if 'spin_kernel' in n]
if len(marks) >= 2:                       # VG_BENCH_MARK=1: the instrumented serial step sits between two marker kernels
    step = rows[marks[-2] + 1: marks[-1]]
else:
    ad = [i for i, (n, _) in enumerate(rows) if 'adamw' in n]
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    step = rows[ad[2 * k - 1] + 1: ad[2 * k + 1] + 1] if len(ad) >= 2 * k + 2 else rows[ad[-3] + 1: ad[-1] + 1]
tot = sum(v for _, v in step)
print(f'launches {len(step)}  total {tot / 1e6:.3f} ms (cold-cache, serialised)')


def short(n):
    n = re.sub(r'^void ', '', n)
    m = re.match(r'([\w:]+)', n)
    base = m.group(1) if m else n[:40]
    if 'elementwise' in base or 'reduce_kernel' in base:
        m2 = re.search(r'(\w+Functor|\w+_kernel_cuda|\w+Ops|func_wrapper\w*|\w+_impl)', n[len(base):])
        if m2:
            base += '<' + m2.group(1) + '>'
    if 'gemm_tc_kernel' in base:
        m2 = re.search(r'gemm_tc_kernel<([^>]*)>', n)
        if m2:
            base += '<' + m2.group(1).replace('(bool)', '') + '>'
    return base[:100]


agg = collections.defaultdict(lambda: [0, 0.0])
for n, v in step:
    a = agg[short(n)]
    a[0] += 1
    a[1] += v
print('| share | time (us) | launches | kernel |\n|---:|---:|---:|---|')
for kname, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:50]:
    print(f'| {v / tot * 100:.2f}% | {v / 1e3:.1f} | {c} | `{kname}` |')
