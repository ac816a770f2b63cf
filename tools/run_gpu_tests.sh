# GPU-box test + measurement driver (run under gpurun).  Every stage has its own timeout so that a hung kernel
# cannot eat the whole lease.
mkdir -p gpurun_out
timeout 120 python tools/debug_decode.py > gpurun_out/debug_decode.log 2>&1; echo "rc=$?" >> gpurun_out/debug_decode.log
timeout 300 python -m pytest tests/test_kernels_gpu.py -q --tb=short > gpurun_out/pytest_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_kernels.log
timeout 500 python -m pytest tests/test_model_gpu.py -q --tb=short > gpurun_out/pytest_model.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_model.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
cat gpurun_out/debug_decode.log; tail -3 gpurun_out/pytest_kernels.log; tail -8 gpurun_out/pytest_model.log; tail -4 gpurun_out/smoke.log; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['roofline'], d['decode'])"; tail -3 gpurun_out/bench.err
