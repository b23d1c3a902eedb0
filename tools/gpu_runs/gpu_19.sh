RAPT_B200_ADAPTIVE_SLICES=1 python tools/bench_configs.py adaptive 1048576 300 fast 1 | cut -c1-200
RAPT_B200_ADAPTIVE_SLICES=8 python tools/bench_configs.py adaptive 1048576 300 fast 1 | cut -c1-200
RAPT_B200_ADAPTIVE_SLICES=16 python tools/bench_configs.py adaptive 1048576 300 fast 1 | cut -c1-200
RAPT_B200_ADAPTIVE_SLICES=32 python tools/bench_configs.py adaptive 1048576 300 fast 1 | cut -c1-200
