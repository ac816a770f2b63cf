mkdir -p gpurun_out
for cfg in "2 3000" "8 640"; do set -- $cfg
timeout 600 python bench.py --batch $1 --frames $2 --steps 10 --warmup 3 --no-cpu-baseline --no-decode > gpurun_out/bench_b$1_t$2.json 2> gpurun_out/bench_b$1_t$2.err; echo "rc=$?" >> gpurun_out/bench_b$1_t$2.err
python -c "
import json
d=json.load(open('gpurun_out/bench_b$1_t$2.json')); print('B=$1 T=$2', d['value'], d['ms_per_step'], d['step_mode'], d['mfu_vs_measured_sustained'], d['roofline']['achieved'], d['loss'])"; tail -1 gpurun_out/bench_b$1_t$2.err
done
