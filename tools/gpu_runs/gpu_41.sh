# bounce-centre path (N4) first GPU run: new parity tests with measured errors, then the whole GPU suite, then a timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bc.py -m gpu -q -rA > gpurun_out/bc_tests.log 2>&1; echo "bc tests exit $?"
grep -E "passed|failed|^(bc_|fast|strict|member|PASSED|FAILED|ERROR)|^E  " gpurun_out/bc_tests.log | cut -c1-400 | tail -80
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_bc.py > gpurun_out/gpu_suite.log 2>&1; echo "suite exit $?"
tail -25 gpurun_out/gpu_suite.log | cut -c1-300
timeout 600 python tools/bench_bc.py 16384 1.0 fast 2>&1 | tail -1 | tee gpurun_out/bench_bc_fast.json | cut -c1-600
timeout 600 python tools/bench_bc.py 4096 1.0 strict 2>&1 | tail -1 | tee gpurun_out/bench_bc_strict.json | cut -c1-600
