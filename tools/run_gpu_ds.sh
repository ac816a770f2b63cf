mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm or rmsnorm or decode or cache" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_parity_full_gpu.py -q -x -k "decode or step or generation or sampler or cached" 2>&1 | tail -3
timeout 600 python tools/decode_bench.py 64 128 256 --kind=layerwise 2>&1 | grep -v Warning | tee gpurun_out/ds_bench_layerwise_pdl.log
VG_DECODE_PDL=0 timeout 600 python tools/decode_bench.py 64 128 256 --kind=layerwise 2>&1 | grep -v Warning | tee gpurun_out/ds_bench_layerwise_nopdl.log
timeout 600 python tools/decode_bench.py 32 64 96 --kind=linear 2>&1 | grep -v Warning | tee gpurun_out/ds_bench_linear.log
