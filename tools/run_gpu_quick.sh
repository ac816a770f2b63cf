mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q --tb=short -x > gpurun_out/pytest_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_kernels.log
timeout 400 python -m pytest tests/test_model_gpu.py -q --tb=short > gpurun_out/pytest_model.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_model.log
timeout 300 python tools/gemm_bench.py --quick > gpurun_out/gemm_bench.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-decode > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
tail -15 gpurun_out/pytest_kernels.log; tail -15 gpurun_out/pytest_model.log; grep -E "ffn|qkv|out_proj" gpurun_out/gemm_bench.log; python -c "
import json
d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['step_mode'], d['e2e']['value'], d['roofline'])"; tail -2 gpurun_out/bench.err
