mkdir -p gpurun_out
( time timeout 185 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_all.log 2>&1; grep "passed\|failed\|Error" gpurun_out/pytest_gpu_all.log | head -5
( timeout 60 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 40 python tools/decode_bench.py 128 256 --kind=layerwise 2>/dev/null | cut -c1-120
