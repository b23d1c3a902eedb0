python -m pytest tests/test_gpu_adaptive.py tests/test_gpu_particle.py tests/test_gpu_gc.py -m gpu -q 2>&1 | tail -5
RAPT_B200_TRACE=1 python tools/bench_configs.py adaptive 65536 300 fast 2 2>gpurun_out/adaptive_trace2.log | cut -c1-300
head -24 gpurun_out/adaptive_trace2.log
RAPT_B200_ADAPTIVE_SLICES=1 python tools/bench_configs.py adaptive 65536 300 fast 1 | cut -c1-200
RAPT_B200_ADAPTIVE_SLICES=64 python tools/bench_configs.py adaptive 65536 300 fast 1 | cut -c1-200
