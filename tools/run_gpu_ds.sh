mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm" 2>&1 | tail -5
for t in 1 0; do echo "== M=5120 VG_GEMM_TAIL=$t"; VG_GEMM_TAIL=$t timeout 300 python tools/gemm_bench.py 5120 --quick 2>&1 | grep "fwd\|dgrad" | awk '{print $1,$2,$3,$5,$6,$7,$8}' | head -8; done 2>&1 | tee gpurun_out/gemm_tail.log
