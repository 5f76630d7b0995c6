timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "optional_backward" 2>&1 | tail -15
