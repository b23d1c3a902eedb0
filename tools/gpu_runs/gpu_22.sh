python -m pytest tests/test_gpu_gc.py -m gpu -q -k bounce_period_device 2>&1 | tail -2
python tools/quick_bench.py 262144 10.0 fast 2 0 1
python tools/quick_bench.py 262144 10.0 fast 2 1 1 12000
python tools/quick_bench.py 262144 10.0 fast 2 8 1 1500
