mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 150 $TR tools/dp_check.py > gpurun_out/dp_check.log 2>&1; echo "dp_check rc=$?"
grep -E "dp_check|Error|assert" gpurun_out/dp_check.log | head
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-decode --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
tail -c 1500 gpurun_out/bench_n2.json
