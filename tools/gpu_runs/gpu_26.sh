python bench.py --steps 3 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench_r1_n1_v7.json | cut -c1-300
python bench.py --impl reference --steps 2 --warmup 1 | tee gpurun_out/bench_r1_ref_v7.json | cut -c1-200
python tools/bench_configs.py gc 1048576 10.0 fast 2 | tee gpurun_out/cfg3_gc_v4.json | cut -c1-300
python tools/bench_configs.py belt 1048576 2.0 fast 2 | tee gpurun_out/cfg5_belt_v4.json | cut -c1-300
python tools/bench_configs.py adaptive 1048576 300 fast 1 | tee gpurun_out/cfg4_adaptive_v4.json | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_v7.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
