timeout 400 python -m pytest tests/test_zz_shared_device_gpu.py -m gpu -q -x 2>&1 | tail -3
