python -m pytest tests/test_gpu_codes.py -m gpu -x -q 2>&1 | tail -15
