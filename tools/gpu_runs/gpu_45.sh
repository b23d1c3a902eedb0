timeout 900 python -m pytest tests/test_gpu_quad.py -m gpu -q -rA -k "curve_integrals and 61-72" > gpurun_out/quad_dev3.log 2>&1; echo "exit $?"
grep -E "^ ex" gpurun_out/quad_dev3.log | cut -c1-300 | head -40
