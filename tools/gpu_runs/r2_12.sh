# Round 2, GPU call 12 (1 GPU): where the end of the particle kernel goes at 1 M tracers (profiling build: per-tracer
# fetch / retire times), suite on the current library.
mkdir -p gpurun_out
RAPT_B200_LIB=$PWD/rapt_b200/librapt_b200_trace.so python tools/tail_profile.py 1048576 > gpurun_out/r2_12_tail_1M.json 2>gpurun_out/r2_12_err.log; cut -c1-2500 gpurun_out/r2_12_tail_1M.json
RAPT_B200_LIB=$PWD/rapt_b200/librapt_b200_trace.so python tools/tail_profile.py 4194304 > gpurun_out/r2_12_tail_4M.json 2>>gpurun_out/r2_12_err.log; cut -c1-1200 gpurun_out/r2_12_tail_4M.json
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_12_pytest.log 2>&1; tail -4 gpurun_out/r2_12_pytest.log | cut -c1-300
tail -3 gpurun_out/r2_12_err.log | cut -c1-300
