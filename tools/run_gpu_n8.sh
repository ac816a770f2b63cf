mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
run() { name=$1; shift; env "$@" timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 3 --no-decode --no-shapes --no-cpu-baseline --no-gpu-reference > gpurun_out/n8_$name.json 2> gpurun_out/n8_$name.err; python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/n8_$name.json').read().strip().splitlines()[-1]); print('$name', d['value'], d['ms_per_step'], d.get('dp_check',{}).get('params_identical_across_ranks'))
except Exception as e: print('$name failed', e)"; grep -m2 -i "nvls\|error" gpurun_out/n8_$name.err | cut -c1-160; }
run default A=1
run nvls NCCL_ALGO=NVLS NCCL_DEBUG=INFO
run chan8 NCCL_MAX_NCHANNELS=8
run chan4 NCCL_MAX_NCHANNELS=4
run bf16 VG_DP_BF16=1
