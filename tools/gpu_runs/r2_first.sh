# First GPU call of round 2: A/B of the particle kernel v9 (h-scaled stage vectors, the default build) against v8
# (-DRAPT_RKN_HK=0), then the evidence refresh.  Build the v8 library HERE (CPU container) before sending:
#   make -C rapt_b200/csrc -j8 B=build_v8 OUT=../librapt_b200_v8.so EXTRA=-DRAPT_RKN_HK=0
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench_r2_n1_v9.json | cut -c1-300
RAPT_B200_LIB=$PWD/rapt_b200/librapt_b200_v8.so python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench_err.log | tee gpurun_out/bench_r2_n1_v8.json | cut -c1-300
python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/bench_err.log | tee gpurun_out/bench_r2_ref.json | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_stdout.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_particle_rkn -c 1 -o gpurun_out/r2_particle_v9 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/ncu_stdout.log 2>&1
tail -3 gpurun_out/bench_err.log | cut -c1-200
# guiding-centre kernel (changes after the last round-1 GPU call: folded coefficients, reciprocals instead of divisions)
python bench.py --workload gc --steps 2 --warmup 2 --no-cpu-baseline 2>>gpurun_out/bench_err.log | tee gpurun_out/bench_r2_gc.json | cut -c1-200
python bench.py --workload belt --steps 2 --warmup 2 --no-cpu-baseline 2>>gpurun_out/bench_err.log | tee gpurun_out/bench_r2_belt.json | cut -c1-200
