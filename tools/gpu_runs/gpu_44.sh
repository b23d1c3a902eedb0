timeout 900 python -m pytest tests/test_gpu_quad.py -m gpu -q -rA -k "curve_integrals and 61-72" > gpurun_out/quad_dev2.log 2>&1; echo "exit $?"
grep -E -A40 "^what 0" gpurun_out/quad_dev2.log | cut -c1-260 | head -70
