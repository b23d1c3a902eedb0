python -m pytest tests -m gpu -q -x 2>&1 | tail -30
python tools/quick_bench.py 1048576 1.0 fast 3 0 1
python tools/quick_bench.py 1048576 10.0 fast 2 0 1
