python -m pytest tests -m gpu -q 2>&1 | tail -12
python bench.py --steps 3 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench_r1_n1_v6.json | cut -c1-400
grep e2e gpurun_out/bench_err.log
ncu --set full --clock-control none --import-source on -k regex:k_particle_rkn -c 1 -o gpurun_out/prof_particle_r1f python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_stdout.log 2>&1
