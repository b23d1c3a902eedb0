python tools/bench_configs.py gc 1048576 10.0 fast 2 | tee gpurun_out/cfg3_gc.json
python tools/bench_configs.py belt 1048576 2.0 fast 2 | tee gpurun_out/cfg5_belt.json
python tools/bench_configs.py adaptive 65536 300 fast 2 | tee gpurun_out/cfg4_adaptive.json
ncu --set full --clock-control none --import-source on -k regex:k_gc_dopri5 -c 1 -o gpurun_out/prof_gc_r1a python tools/bench_configs.py gc 262144 2.0 fast 1 > gpurun_out/ncu_stdout_gc.log 2>&1
tail -2 gpurun_out/ncu_stdout_gc.log | cut -c1-300
