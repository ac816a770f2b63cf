mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"
tail -3 gpurun_out/bench_n1.err
python - <<'P'
import json
d = json.loads([l for l in open("gpurun_out/bench_n1.json") if l.startswith("{")][-1])
for k in ("value", "ms_per_step", "mfu_vs_measured_sustained", "mfu_vs_nominal_2250", "e2e", "roofline", "shapes", "cpu_baseline", "gpu_reference", "gpu_launches", "clocks"):
    print(k, json.dumps(d.get(k)))
for k, v in d.get("decode", {}).items():
    print(k, json.dumps(v))
P
