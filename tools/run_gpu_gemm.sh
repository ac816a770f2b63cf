mkdir -p gpurun_out
# CTA-pair kernels first, under a short timeout (a protocol bug would hang)
timeout 180 python -m pytest tests/test_kernels_gpu.py -q --tb=short -x -k "gemm" > gpurun_out/pytest_gemm.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gemm.log
tail -8 gpurun_out/pytest_gemm.log
if grep -q "rc=0" gpurun_out/pytest_gemm.log; then
timeout 200 python tools/gemm_bench.py --quick > gpurun_out/gemm_bench.log 2>&1
cat gpurun_out/gemm_bench.log
VG_GEMM_PAIR=0 timeout 200 python tools/gemm_bench.py --quick > gpurun_out/gemm_bench_nopair.log 2>&1
echo "=== VG_GEMM_PAIR=0"; cat gpurun_out/gemm_bench_nopair.log
fi
