python -m pytest tests/test_gpu_bc.py -m gpu -q 2>&1 | tail -12 | cut -c1-300
