mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "decode or cache" 2>&1 | tail -3
timeout 600 python tools/decode_bench.py 64 128 256 --kind=layerwise 2>&1 | grep -v Warning | tee gpurun_out/ds_bench_layerwise.log
timeout 600 python tools/decode_bench.py 32 64 96 --kind=linear 2>&1 | grep -v Warning | tee gpurun_out/ds_bench_linear.log
