python -m pytest tests -m gpu -q -x 2>&1 | tail -8
python tools/quick_bench.py 1048576 10.0 fast 2 0 1
RAPT_B200_GC_BLOCKS=2 python tools/bench_configs.py gc 1048576 10.0 fast 2 | cut -c1-400
RAPT_B200_GC_BLOCKS=3 python tools/bench_configs.py gc 1048576 10.0 fast 2 | cut -c1-400
RAPT_B200_GC_BLOCKS=2 python tools/bench_configs.py belt 1048576 2.0 fast 2 | cut -c1-400
RAPT_B200_GC_BLOCKS=3 python tools/bench_configs.py belt 1048576 2.0 fast 2 | cut -c1-400
RAPT_B200_GC_BLOCKS=3 ncu --set full --clock-control none --import-source on -k regex:k_gc_dopri5 -c 1 -o gpurun_out/prof_gc_r1b python tools/bench_configs.py gc 262144 2.0 fast 1 > gpurun_out/ncu_stdout_gc.log 2>&1
