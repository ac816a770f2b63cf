for i in 1 2; do timeout 900 python bench.py --no-decode --no-shapes --no-gpu-reference --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print(d['value'], d['ms_per_step'], r['achieved'], r['frac'], r['gemm_ms_per_step'], d['clocks'])"; done
