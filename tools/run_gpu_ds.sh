mkdir -p gpurun_out
timeout 125 python bench.py --no-gpu-reference --no-cpu-baseline --no-shapes --steps 10 > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_last.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'])
for k,v in d['decode'].items():
    print(k, v.get('ms_per_step'), v.get('frac'), v.get('path', v.get('mode','')))
PY
