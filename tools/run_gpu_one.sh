mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_model_gpu.py -q --tb=short -k "train_step" 2>&1 | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-decode > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
python -c "
import json
d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['step_mode'], d['e2e']['value'], d['roofline']['gemm_ms_per_step'], d['gpu_launches'], d['loss'])"; tail -2 gpurun_out/bench.err
