mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_model_gpu.py -q --tb=short -x -k "train_step or first_writer or arena" > gpurun_out/pytest_arena.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_arena.log
tail -15 gpurun_out/pytest_arena.log
timeout 400 python bench.py --no-cpu-baseline --no-decode > gpurun_out/bench_mid.json 2> gpurun_out/bench_mid.err; echo "rc=$?" >> gpurun_out/bench_mid.err
python -c "
import json
d=json.load(open('gpurun_out/bench_mid.json')); print(d['value'], d['ms_per_step'], d['mfu_vs_measured_sustained'], d['step_mode'], d['e2e']['value'], d['roofline']['gemm_ms_per_step'], d['roofline']['achieved'], d['loss'], d['clocks'])"; tail -2 gpurun_out/bench_mid.err
