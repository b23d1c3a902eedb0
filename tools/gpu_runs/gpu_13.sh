python -m pytest tests -m gpu -q 2>&1 | tail -6
python bench.py --steps 3 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench_r1_n1_v5.json | cut -c1-700
ncu --set full --clock-control none --import-source on -k regex:k_particle_rkn -c 1 -o gpurun_out/prof_particle_r1e python tools/quick_bench.py 1048576 1.0 fast 1 0 1 > gpurun_out/ncu_stdout.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_v5.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
