timeout 300 python -m pytest tests/test_model_gpu.py -q --tb=short -k "ddim or likelihood" 2>&1 | tail -15
