# guiding-centre kernel: cooperative HINIT probe (COOP) -- parity tests, then config 3 / config 5 with and without it
timeout 1200 python -m pytest tests/test_gpu_gc.py tests/test_gpu_adaptive.py tests/test_gpu_properties.py tests/test_gpu_userfield.py tests/test_gpu_grid.py -m gpu -q 2>&1 | tail -8 | cut -c1-300
for c in 0 1; do
RAPT_B200_GC_COOP=$c python bench.py --workload gc --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>>gpurun_out/bench_err.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('gc coop=$c', d['value'], d['ms_per_step'], d['roofline']['frac'], d['config']['particle_steps_per_bench_step'])"
RAPT_B200_GC_COOP=$c python bench.py --workload belt --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>>gpurun_out/bench_err.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('belt coop=$c', d['value'], d['ms_per_step'], d['roofline']['frac'], d['config']['particle_steps_per_bench_step'])"
done
tail -3 gpurun_out/bench_err.log
