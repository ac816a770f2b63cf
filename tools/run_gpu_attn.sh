mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q --tb=short -x -k "attention" > gpurun_out/pytest_attn.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_attn.log
tail -3 gpurun_out/pytest_attn.log
timeout 200 python tools/attn_bench.py > gpurun_out/attn_bench.log 2>&1; grep tcgen05 gpurun_out/attn_bench.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-decode > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
python -c "
import json
d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['step_mode'], d['e2e']['value'], d['roofline']['gemm_ms_per_step'], d['roofline']['achieved'], d['loss'])"; tail -2 gpurun_out/bench.err
