python -m pytest tests -m gpu -q 2>&1 | tail -25
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --steps 3 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench_r1_n1.json
tail -5 gpurun_out/bench_err.log
python bench.py --impl reference --steps 2 --warmup 1 | tee gpurun_out/bench_r1_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/bench_under_ncu.log | cut -c1-300
