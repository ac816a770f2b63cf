mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention_decode" 2>&1 | tail -2
VG_AD_CFG=128382 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention_decode" 2>&1 | tail -2
for cfg in 64342 64343 64344 64244 64542 128382 64682 64482 32882; do
echo "== VG_AD_CFG=$cfg"
VG_AD_CFG=$cfg timeout 200 python tools/attn_decode_bench.py 402 8 64 256 2>&1 | grep "splits=1\|splits=2"
done 2>&1 | tee gpurun_out/ad_bench_cfg2.log
