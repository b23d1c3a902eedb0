python -m pytest tests -m gpu -q 2>&1 | tail -8
python tools/bench_configs.py gc 1048576 10.0 fast 2 | cut -c1-330
python tools/bench_configs.py belt 1048576 2.0 fast 2 | cut -c1-330
