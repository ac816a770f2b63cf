timeout 900 python -m pytest tests/test_decode_step_gpu.py -q -x 2>&1 | tail -2
echo "== prefetch on"; timeout 600 python tools/decode_bench.py 1 4 8 16 --kind=step 2>&1 | grep -v Warning | cut -c1-80
echo "== prefetch off"; VG_DS_ATTN_PREFETCH=0 timeout 600 python tools/decode_bench.py 1 4 8 16 --kind=step 2>&1 | grep -v Warning | cut -c1-80
