mkdir -p gpurun_out
VG_GEMM_PAIR128=2 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm" 2>&1 | tail -3
for M in 5120 8000 6000; do for p in 0 2; do
echo "== M=$M VG_GEMM_PAIR128=$p"; VG_GEMM_PAIR128=$p timeout 300 python tools/gemm_bench.py $M --quick 2>&1 | grep "fwd\|dgrad" | awk '{print $1,$2,$3,$5,$6,$7,$8}'
done; done 2>&1 | tee gpurun_out/gemm_pair128.log
