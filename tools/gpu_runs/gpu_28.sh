# separable VarEarthDipole in the GC fast path: parity + belt bench
python -m pytest tests/test_gpu_gc.py tests/test_gpu_adaptive.py tests/test_gpu_properties.py -q -m gpu -x 2>&1 | tail -5
python bench.py --workload belt --steps 2 --warmup 1 --no-e2e --cpu-sample 512 2>gpurun_out/belt_sep_err.log | tee gpurun_out/bench_r1_belt_n1_sep.json | cut -c1-900
tail -2 gpurun_out/belt_sep_err.log
