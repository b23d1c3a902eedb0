RAPT_B200_NO_RKN=1 python tools/bench_grid.py 1048576 1.0 1 129 fast 2>&1 | tail -1 | cut -c1-420
