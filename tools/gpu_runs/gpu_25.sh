timeout 300 python -m pytest tests/test_gpu_properties.py -m gpu -q 2>&1 | tail -8
timeout 300 python tools/quick_bench.py 1048576 10.0 fast 2 0 1
