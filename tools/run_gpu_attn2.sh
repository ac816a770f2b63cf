mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_kernels_gpu.py -q --tb=short -x -k "attention" > gpurun_out/pytest_attn.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_attn.log
tail -3 gpurun_out/pytest_attn.log
timeout 60 python tools/attn_trace.py 8 1000 16 0 900 > gpurun_out/attn_trace.log 2>&1; cat gpurun_out/attn_trace.log | cut -c1-150
timeout 100 python tools/attn_raw.py > gpurun_out/attn_raw.log 2>&1; cat gpurun_out/attn_raw.log
