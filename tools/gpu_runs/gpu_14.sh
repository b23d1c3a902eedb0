python -m pytest tests/test_gpu_gc.py -m gpu -q -k pa90 2>&1 | tail -3
python tools/e2e_probe.py
for L in librapt_b200.so librapt_b200_nopolish.so librapt_b200_fence.so; do echo $L; RAPT_B200_LIB=/root/repo/rapt_b200/$L python tools/quick_bench.py 1048576 10.0 fast 2 0 1; done
RAPT_B200_LIB=/root/repo/rapt_b200/librapt_b200_nopolish.so python -m pytest tests/test_gpu_particle.py tests/test_gpu_gc.py tests/test_gpu_properties.py -m gpu -q 2>&1 | tail -4
RAPT_B200_LIB=/root/repo/rapt_b200/librapt_b200_nopolish.so python tools/bench_configs.py gc 1048576 10.0 fast 2 | cut -c1-300
