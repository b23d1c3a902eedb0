# bounce-centre: tests after the tolerance fixes + user-field test, lanes per SM, ncu capture of k_bounce_center
timeout 900 python -m pytest tests/test_gpu_bc.py tests/test_gpu_quad.py -m gpu -q -rA > gpurun_out/bc_tests3.log 2>&1; echo "bc+quad tests exit $?"
grep -E "passed|failed|^(status|FAILED|ERROR)|^E  " gpurun_out/bc_tests3.log | cut -c1-300 | tail -20
RAPT_B200_BC_BLOCKS=4 timeout 600 python tools/bench_bc.py 65536 1.0 fast 2>&1 | tail -1 | cut -c1-420
RAPT_B200_BC_BLOCKS=8 timeout 600 python tools/bench_bc.py 65536 1.0 fast 2>&1 | tail -1 | tee gpurun_out/bench_bc_65536.json | cut -c1-420
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bounce_center -c 1 -o gpurun_out/prof_bc_v1 python tools/bench_bc.py 16384 0.05 fast > gpurun_out/ncu_bc.log 2>&1; tail -2 gpurun_out/ncu_bc.log | cut -c1-200
