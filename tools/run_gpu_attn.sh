mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q --tb=short -x -k "attention" > gpurun_out/pytest_attn.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_attn.log
tail -6 gpurun_out/pytest_attn.log
timeout 200 python tools/attn_bench.py > gpurun_out/attn_bench.log 2>&1; cat gpurun_out/attn_bench.log
