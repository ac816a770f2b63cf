mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --tb=short -x > gpurun_out/pytest_gpu_all.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_all.log
tail -4 gpurun_out/pytest_gpu_all.log
timeout 400 python bench.py --no-cpu-baseline --no-decode > gpurun_out/bench_mid.json 2> gpurun_out/bench_mid.err; echo "rc=$?" >> gpurun_out/bench_mid.err
cat gpurun_out/bench_mid.json; tail -2 gpurun_out/bench_mid.err
