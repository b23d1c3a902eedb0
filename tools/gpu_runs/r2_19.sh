# Round 2, GPU call 19 (1 GPU): the rest of the suite after the fix of test_launch_shapes_agree (it compared unwritten rows).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_19_pytest.log 2>&1; tail -15 gpurun_out/r2_19_pytest.log | cut -c1-300
