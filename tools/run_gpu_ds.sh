mkdir -p gpurun_out
timeout 600 python tools/decode_bench.py 64 72 80 96 --kind=layerwise 2>&1 | grep -v Warning | tee gpurun_out/ds_bench_layerwise_mid.log
timeout 600 python tools/decode_bench.py 48 64 72 80 --kind=linear 2>&1 | grep -v Warning | tee gpurun_out/ds_bench_linear_mid.log
timeout 600 python tools/decode_bench.py 32 48 --kind=step 2>&1 | grep -v Warning | tee gpurun_out/ds_bench_step_mid.log
VG_BENCH_GEMM_TABLE=1 timeout 900 python bench.py --no-decode --no-shapes --no-gpu-reference --no-cpu-baseline > gpurun_out/bench_train_only.json 2> gpurun_out/bench_train_only.err
tail -1 gpurun_out/bench_train_only.json | cut -c1-1500; head -30 gpurun_out/bench_train_only.err
