python -m pytest tests/test_gpu_particle.py tests/test_gpu_gc.py -m gpu -q 2>&1 | tail -15
python tools/quick_bench.py 1048576 1.0 fast 3 0 0
python tools/quick_bench.py 1048576 1.0 fast 3 0 1
python tools/quick_bench.py 1048576 10.0 fast 2 0 1
ncu --set full --clock-control none --import-source on -k regex:k_particle_dop853 -c 1 -o gpurun_out/prof_particle_r1c python tools/quick_bench.py 262144 0.5 fast 1 0 1 > gpurun_out/ncu_stdout.log 2>&1
tail -3 gpurun_out/ncu_stdout.log
