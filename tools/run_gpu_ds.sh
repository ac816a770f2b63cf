mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py -q -x -k "utterance_encoder" 2>&1 | tail -8
