python -m pytest tests/test_gpu_grid.py -q -m gpu 2>&1 | tail -4
echo "--- RKN + cell cache"; python tools/bench_grid.py 1048576 1.0 1 129 fast 2>&1 | tail -1 | cut -c1-330
echo "--- 2 time points"; python tools/bench_grid.py 1048576 1.0 2 129 fast 2>&1 | tail -1 | cut -c1-330
echo "--- 257^2x129"; python tools/bench_grid.py 1048576 1.0 1 257 fast 2>&1 | tail -1 | cut -c1-330
