timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 100 -k "float64_integer_kinds" 2>&1 | tail -6
