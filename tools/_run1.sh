(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_slab.py -m gpu -x -q 2>&1 | tail -3)
