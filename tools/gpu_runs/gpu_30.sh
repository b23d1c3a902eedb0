ncu --set full --clock-control none --import-source on -k regex:k_particle_rkn -c 1 -o gpurun_out/prof_particle_r1g python tools/quick_bench.py 1048576 0.5 fast 1 > gpurun_out/ncu_stdout.log 2>&1
tail -3 gpurun_out/ncu_stdout.log
