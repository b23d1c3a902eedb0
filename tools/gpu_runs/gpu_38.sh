ncu --set full --clock-control none --import-source on -k regex:k_particle_rkn -c 1 -o gpurun_out/prof_grid_r1b python tools/bench_grid.py 262144 0.5 1 129 fast > gpurun_out/ncu_grid.log 2>&1
tail -1 gpurun_out/ncu_grid.log | cut -c1-200
