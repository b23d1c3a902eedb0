# Round 2, GPU call 1: the whole GPU suite on the round-1 head (no -x, so every failure is listed), then the A/B of
# particle kernel v9 (default) against v8 (-DRAPT_RKN_HK=0, built beforehand as librapt_b200_v8.so), the guiding-centre
# lines, and the ncu evidence of the benchmarked binary.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_01_pytest.log 2>&1; tail -15 gpurun_out/r2_01_pytest.log
python bench.py --steps 3 --warmup 3 2>gpurun_out/bench_err.log > gpurun_out/r2_01_bench_n1_v9.json; cut -c1-400 gpurun_out/r2_01_bench_n1_v9.json
RAPT_B200_LIB=$PWD/rapt_b200/librapt_b200_v8.so python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>>gpurun_out/bench_err.log > gpurun_out/r2_01_bench_n1_v8.json; cut -c1-200 gpurun_out/r2_01_bench_n1_v8.json
python bench.py --workload gc --steps 2 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench_err.log > gpurun_out/r2_01_bench_gc.json; cut -c1-200 gpurun_out/r2_01_bench_gc.json
python bench.py --workload belt --steps 2 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench_err.log > gpurun_out/r2_01_bench_belt.json; cut -c1-200 gpurun_out/r2_01_bench_belt.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_01_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_stdout.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_particle_rkn -c 1 -o gpurun_out/r2_01_particle_v9 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/ncu_stdout.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_gc_dopri5 -c 1 -o gpurun_out/r2_01_gc python bench.py --workload gc --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/ncu_stdout2.log 2>&1
tail -3 gpurun_out/bench_err.log | cut -c1-200
