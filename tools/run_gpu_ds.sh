timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "skinny" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_parity_full_gpu.py -q -x -k "cached_generation" 2>&1 | tail -2
echo "== mixed"; timeout 600 python tools/decode_bench.py 72 96 128 --kind=layerwise 2>&1 | grep -v Warning | cut -c1-100
echo "== narrow off"; VG_SKINNY_MAX_ROWS_NARROW=64 timeout 600 python tools/decode_bench.py 72 96 128 --kind=layerwise 2>&1 | grep -v Warning | cut -c1-100
