mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_kernels_gpu.py -q --tb=short -x -k "attention" > gpurun_out/pytest_attn.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_attn.log
tail -3 gpurun_out/pytest_attn.log
timeout 60 python tools/attn_trace.py > gpurun_out/attn_trace.log 2>&1; cat gpurun_out/attn_trace.log
timeout 100 python tools/attn_bench.py > gpurun_out/attn_bench.log 2>&1; grep tcgen05 gpurun_out/attn_bench.log
