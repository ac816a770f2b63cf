mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q --tb=short -x -k "decode" > gpurun_out/pytest_dl.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_dl.log
tail -3 gpurun_out/pytest_dl.log
timeout 400 python -m pytest tests/test_model_gpu.py -q --tb=short -k "decode" > gpurun_out/pytest_model_dec.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_model_dec.log
tail -3 gpurun_out/pytest_model_dec.log
timeout 600 python tools/decode_bench.py 1 8 32 64 > gpurun_out/decode_bench.log 2>&1; echo "rc=$?" >> gpurun_out/decode_bench.log
tail -3 gpurun_out/decode_bench.log
