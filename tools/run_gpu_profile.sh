# Full default bench (decode + cpu baseline), then the ncu launch list of an eager training step.
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "rc=$?" >> gpurun_out/bench_full.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/launches_eager.csv \
  python bench.py --steps 1 --warmup 1 --no-cuda-graph --no-cpu-baseline --no-decode > gpurun_out/ncu_eager.log 2>&1; echo "rc=$?" >> gpurun_out/ncu_eager.log
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_full.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('cpu_baseline'), d.get('decode'))
PY
tail -3 gpurun_out/bench_full.err; tail -3 gpurun_out/ncu_eager.log; wc -l gpurun_out/launches_eager.csv
