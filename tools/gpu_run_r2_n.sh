python -m pytest tests/test_gpu_fuzz.py -q -x 2>&1 | tail -40
