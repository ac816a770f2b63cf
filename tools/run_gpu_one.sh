timeout 300 python -m pytest tests/test_model_gpu.py -q --tb=short -k "train_step" 2>&1 | tail -15
