# round-1 evidence refresh, session 3: whole GPU suite, bench lines (ours, reference arm, gc, belt), launch list, ncu of the particle kernel
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench_r1_n1_s3.json | cut -c1-300
python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/bench_err.log | tee gpurun_out/bench_r1_ref_s3.json | cut -c1-300
python bench.py --workload gc --steps 2 --warmup 2 2>>gpurun_out/bench_err.log | tee gpurun_out/bench_r1_gc_s3.json | cut -c1-200
python bench.py --workload belt --steps 2 --warmup 2 2>>gpurun_out/bench_err.log | tee gpurun_out/bench_r1_belt_s3.json | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_s3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_stdout.log 2>&1
tail -3 gpurun_out/bench_err.log | cut -c1-200
