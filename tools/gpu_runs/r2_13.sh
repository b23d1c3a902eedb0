# Round 2, GPU call 13 (1 GPU): the first wave of the particle kernel spread over the warps (AdvArgs.spread_first_wave) --
# A/B against RAPT_B200_NO_SPREAD=1 in one call, per-tracer fetch/retire times of both, suite on the library.
mkdir -p gpurun_out
run() { # tag env
  env $2 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra --no-e2e 2>>gpurun_out/r2_13_err.log > gpurun_out/r2_13_$1.json
  python -c "
import json; d=json.load(open('gpurun_out/r2_13_$1.json')); print('$1', round(d['ms_per_step'],2), '%.4g'%d['value'], round(d['roofline']['frac'],4))"
}
run spread_a X=1
run nospread_a RAPT_B200_NO_SPREAD=1
run spread_b X=1
run nospread_b RAPT_B200_NO_SPREAD=1
RAPT_B200_LIB=$PWD/rapt_b200/librapt_b200_trace.so python tools/tail_profile.py 1048576 gpurun_out/r2_13_tail_spread.npz > gpurun_out/r2_13_tail_spread.json 2>>gpurun_out/r2_13_err.log; cut -c1-700 gpurun_out/r2_13_tail_spread.json
RAPT_B200_NO_SPREAD=1 RAPT_B200_LIB=$PWD/rapt_b200/librapt_b200_trace.so python tools/tail_profile.py 1048576 gpurun_out/r2_13_tail_nospread.npz > gpurun_out/r2_13_tail_nospread.json 2>>gpurun_out/r2_13_err.log; cut -c1-700 gpurun_out/r2_13_tail_nospread.json
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_13_pytest.log 2>&1; tail -4 gpurun_out/r2_13_pytest.log | cut -c1-300
tail -3 gpurun_out/r2_13_err.log | cut -c1-300
