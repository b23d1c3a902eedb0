python -m pytest tests/test_gpu_userfield.py -m gpu -q 2>&1 | tail -3
python bench.py --workload gc --steps 2 --warmup 2 2>gpurun_out/e.log | tee gpurun_out/bench_r1_gc_n1.json | cut -c1-900
python bench.py --workload belt --steps 2 --warmup 2 2>>gpurun_out/e.log | tee gpurun_out/bench_r1_belt_n1.json | cut -c1-400
tail -3 gpurun_out/e.log
