for p in 0 1 0 1; do VG_DECODE_PDL=$p timeout 300 python tools/ddim_bench.py 2>&1 | grep VG_DECODE_PDL; done
timeout 600 python -m pytest tests/test_model_gpu.py -q -x -k "ddim" 2>&1 | tail -2
