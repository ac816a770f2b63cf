# final-build evidence: GPU tests, smoke(), the bench line.  Outputs under gpurun_out/ (scratch).
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/pytest_gpu_all.log 2>&1; tail -4 gpurun_out/pytest_gpu_all.log
( timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
VG_BENCH_GEMM_TABLE=1 timeout 1500 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_n1.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -c 400 gpurun_out/bench_ref.json
