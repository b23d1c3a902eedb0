python tools/quick_bench.py 1048576 1.0 fast 3
python tools/quick_bench.py 1048576 1.0 strict 2
python tools/quick_bench.py 1048576 1.0 fast 2 16
python tools/quick_bench.py 262144 10.0 fast 2
ncu --set full --clock-control none --import-source on -k regex:k_particle -c 1 -o gpurun_out/prof_particle_r1a python tools/quick_bench.py 262144 0.25 fast 1 > gpurun_out/ncu_stdout.log 2>&1
ls -la gpurun_out/
