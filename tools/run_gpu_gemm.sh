mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q --tb=short -x -k "gemm" > gpurun_out/pytest_gemm.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gemm.log
tail -5 gpurun_out/pytest_gemm.log
timeout 300 python tools/gemm_bench.py --quick > gpurun_out/gemm_bench.log 2>&1
cat gpurun_out/gemm_bench.log
