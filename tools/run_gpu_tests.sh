# the whole GPU suite, as the driver runs it at round end (plus timing per test file)
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/ -q -m gpu --durations=12 ${PYTEST_ARGS} > gpurun_out/pytest_gpu_all.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_all.log
grep -E "bf16 gradient summary|passed|failed|^FAILED|^E  " gpurun_out/pytest_gpu_all.log | head -60
tail -25 gpurun_out/pytest_gpu_all.log
