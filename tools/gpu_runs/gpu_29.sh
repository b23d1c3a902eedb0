# straight-line RKN kernel: parity + timing
python -m pytest tests/test_gpu_particle.py tests/test_gpu_adaptive.py tests/test_gpu_properties.py -q -m gpu -x 2>&1 | tail -5
python tools/quick_bench.py 1048576 10 fast 3 2>&1 | tail -3
