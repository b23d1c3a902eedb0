python -m pytest tests/test_gpu_particle.py tests/test_gpu_adaptive.py tests/test_gpu_properties.py -m gpu -q 2>&1 | tail -3
python tools/quick_bench.py 1048576 10.0 fast 2 0 1
RAPT_B200_TRACE=1 python tools/bench_configs.py adaptive 65536 300 fast 1 2>gpurun_out/adaptive_trace.log | cut -c1-300
head -60 gpurun_out/adaptive_trace.log
