# GPU-box test + measurement driver (run under gpurun).  Every stage has its own timeout so that a hung kernel
# cannot eat the whole lease.
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
timeout 150 python -m pytest tests/test_kernels_gpu.py -q --tb=short -x -k "attention_fwd_bwd and tcgen05" > gpurun_out/pytest_attn_tc.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_attn_tc.log
timeout 300 python -m pytest tests/test_kernels_gpu.py -q --tb=short -k "not (attention_fwd_bwd and tcgen05)" > gpurun_out/pytest_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_kernels.log
timeout 500 python -m pytest tests/test_model_gpu.py -q --tb=short > gpurun_out/pytest_model.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_model.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
timeout 200 python tools/attn_bench.py > gpurun_out/attn_bench.log 2>&1; echo "rc=$?" >> gpurun_out/attn_bench.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-decode > gpurun_out/ncu_bench.log 2>&1
tail -4 gpurun_out/pytest_attn_tc.log; tail -3 gpurun_out/pytest_kernels.log; tail -8 gpurun_out/pytest_model.log; tail -4 gpurun_out/smoke.log; cat gpurun_out/attn_bench.log; head -c 1500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
