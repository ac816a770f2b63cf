mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_kernels_gpu.py -q --tb=short -x > gpurun_out/pytest_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_kernels.log
timeout 400 python -m pytest tests/test_model_gpu.py -q --tb=short > gpurun_out/pytest_model.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_model.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-decode > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
tail -4 gpurun_out/pytest_kernels.log; tail -4 gpurun_out/pytest_model.log; python -c "
import json
d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['step_mode'], d['e2e']['value'], d['roofline']['gemm_ms_per_step'], d['gpu_launches'])"; tail -2 gpurun_out/bench.err
