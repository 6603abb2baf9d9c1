timeout 200 python -m pytest tests/test_c_api.py -m gpu -q 2>&1 | tail -2
