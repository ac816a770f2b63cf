mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-decode > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?" >> gpurun_out/bench_n2.err
python - <<'PY'
import json
for n in (1,2):
    try:
        d=json.loads([l for l in open(f'gpurun_out/bench_n{n}.json') if l.startswith('{')][-1])
        print(n, d['value'], d['ms_per_step'], d['per_gpu_frames_per_sec'], d['e2e']['value'], d['roofline']['gemm_ms_per_step'])
    except Exception as e: print(n, 'ERR', e)
PY
tail -3 gpurun_out/bench_n2.err
