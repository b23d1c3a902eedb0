# device build of rapt_quad.cuh vs scipy / host build (diagnosis of the QUADPACK route on the device)
timeout 900 python -m pytest tests/test_gpu_quad.py -m gpu -q -rA > gpurun_out/quad_dev.log 2>&1; echo "exit $?"
grep -E "passed|failed|^FAILED" gpurun_out/quad_dev.log | tail -40
grep -E -A3 "^what " gpurun_out/quad_dev.log | head -60
