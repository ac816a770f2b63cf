mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_model_gpu.py -q --tb=short -k "cuda_graph or bf16_cached" > gpurun_out/pytest_model.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_model.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-decode --no-cuda-graph > gpurun_out/bench_eager.json 2> gpurun_out/bench_eager.err
tail -6 gpurun_out/pytest_model.log; tail -3 gpurun_out/smoke.log; python -c "
import json
for f in ('bench.json','bench_eager.json'):
    d=json.load(open('gpurun_out/'+f)); print(f, d['value'], d['ms_per_step'], d['step_mode'], d['e2e']['value'], d['roofline']['share_of_step'], d.get('decode'), d['clocks'])"; tail -3 gpurun_out/bench.err
