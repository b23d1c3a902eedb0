# round-1 evidence refresh with the final kernels
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench_r1_n1_final.json | cut -c1-300
python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/bench_err.log | tee gpurun_out/bench_r1_ref_final.json | cut -c1-300
python bench.py --workload gc --steps 2 --warmup 2 2>>gpurun_out/bench_err.log | tee gpurun_out/bench_r1_gc_final.json | cut -c1-200
python bench.py --workload belt --steps 2 --warmup 2 2>>gpurun_out/bench_err.log | tee gpurun_out/bench_r1_belt_final.json | cut -c1-200
python tools/bench_configs.py adaptive 1048576 300 fast 2 2>&1 | tail -2 | cut -c1-600
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_stdout.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_particle_rkn -c 1 -o gpurun_out/prof_particle_final python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/ncu_final.log 2>&1
tail -2 gpurun_out/ncu_final.log | cut -c1-200
