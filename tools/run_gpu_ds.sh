timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "skinny and 3072-1024-200" 2>&1 | grep -B30 "^FAILED\|Error" | head -60
