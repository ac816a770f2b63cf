mkdir -p gpurun_out
timeout 300 python tools/gemm_epi_bench.py --cold 2>&1 | grep -v Warn | tee gpurun_out/gemm_epi_cold.log
timeout 300 python tools/gemm_epi_bench.py 2>&1 | grep -v Warn | tee gpurun_out/gemm_epi_hot.log
