# Round 2, GPU call 7 (1 GPU): guiding-centre kernel with batched HINIT probes (two tracers per lane) -- suite on it, A/B
# against the same source with the batching compiled out (-DRAPT_GC_DEFER=0), adaptive, ncu of the new kernel.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_07_pytest.log 2>&1; tail -6 gpurun_out/r2_07_pytest.log | cut -c1-300
run() { # tag lib workload
  RAPT_B200_LIB=$PWD/rapt_b200/$2 python bench.py --workload $3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>>gpurun_out/r2_07_err.log > gpurun_out/r2_07_$1.json
  python -c "
import json; d=json.load(open('gpurun_out/r2_07_$1.json')); print('$1', round(d['ms_per_step'],2), '%.4g'%d['value'], round(d['roofline']['frac'],4), d['config'].get('kernel_ms'))"
}
run gc_defer_a librapt_b200.so gc
run gc_nodefer_a librapt_b200_nodefer.so gc
run gc_defer_b librapt_b200.so gc
run gc_nodefer_b librapt_b200_nodefer.so gc
run belt_defer_a librapt_b200.so belt
run belt_nodefer_a librapt_b200_nodefer.so belt
run belt_defer_b librapt_b200.so belt
run belt_nodefer_b librapt_b200_nodefer.so belt
run adaptive_defer librapt_b200.so adaptive
run adaptive_nodefer librapt_b200_nodefer.so adaptive
ncu --set full --clock-control none --import-source on -k regex:k_gc_dopri5 -c 1 -o gpurun_out/r2_07_gc python bench.py --workload gc --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/r2_07_ncu.log 2>&1
tail -3 gpurun_out/r2_07_err.log | cut -c1-300
