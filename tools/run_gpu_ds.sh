mkdir -p gpurun_out
VG_BENCH_GEMM_TABLE=1 timeout 1500 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['mfu_vs_measured_sustained'])
print(d['gpu_reference']); print(d['cpu_baseline'])
print({k:(v['value'],v['ms_per_step']) for k,v in d['shapes'].items()})
for k,v in d['decode'].items(): print(k, v.get('ms_per_step'), v.get('frac'), v.get('path'))"
