timeout 900 python -m pytest tests -m gpu -q --timeout 200 2>&1 | tail -8
