python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/e.log | tee gpurun_out/bench_tmp.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'], d['e2e'])"
tail -2 gpurun_out/e.log
nvidia-smi --query-gpu=power.draw,power.limit,clocks.sm,temperature.gpu --format=csv
