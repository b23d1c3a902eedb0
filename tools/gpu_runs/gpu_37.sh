python -m pytest tests/test_gpu_grid.py -q -m gpu 2>&1 | tail -4
echo "--- RKN (noinline grid_eval)"; python tools/bench_grid.py 1048576 1.0 1 129 fast 2>&1 | tail -1 | cut -c1-330
echo "--- generic"; RAPT_B200_NO_RKN=1 python tools/bench_grid.py 1048576 1.0 1 129 fast 2>&1 | tail -1 | cut -c1-330
echo "--- RKN 2 time points"; python tools/bench_grid.py 1048576 1.0 2 129 fast 2>&1 | tail -1 | cut -c1-330
