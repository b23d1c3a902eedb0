python -m pytest tests/test_gpu_gc.py -m gpu -q 2>&1 | tail -3
python tools/bench_configs.py gc 1048576 10.0 fast 2 | tee gpurun_out/cfg3_gc_v3.json | cut -c1-330
ncu --set full --clock-control none --import-source on -k regex:k_gc_dopri5 -c 1 -o gpurun_out/prof_gc_r1c python tools/bench_configs.py gc 1048576 10.0 fast 1 > gpurun_out/ncu_stdout_gc.log 2>&1
