python tools/quick_bench.py 65536 10.0 fast 2 0 1
python tools/quick_bench.py 65536 10.0 fast 2 1 1 12000
