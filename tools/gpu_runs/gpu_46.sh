# QUADPACK routines as real device functions: device-vs-scipy quadrature tests, bounce-centre parity, the tests that failed in gpu_41
timeout 900 python -m pytest tests/test_gpu_quad.py -m gpu -q 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_bc.py -m gpu -q -rA > gpurun_out/bc_tests2.log 2>&1; echo "bc tests exit $?"
grep -E "passed|failed|^(bc_|fast|strict|member|FAILED|ERROR)|^E  " gpurun_out/bc_tests2.log | cut -c1-330 | tail -60
timeout 900 python -m pytest tests/test_gpu_adaptive.py tests/test_gpu_gc.py tests/test_gpu_userfield.py -m gpu -q 2>&1 | tail -12 | cut -c1-300
