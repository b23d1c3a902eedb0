# Round 2, GPU call 15 (1 GPU): one 512-thread block per SM as the default launch shape of the particle kernel -- suite on it,
# A/B against 2 x 256 and 4 x 128; the same question for the guiding-centre kernel (default 4 x 128 vs 1 x 512 vs 2 x 256)
# on configs 3, 5 and 4; per-tracer time profile of the new default; the full default bench line.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_15_pytest.log 2>&1; tail -4 gpurun_out/r2_15_pytest.log | cut -c1-300
run() { # tag lib workload
  RAPT_B200_LIB=$PWD/rapt_b200/$2 python bench.py --workload $3 --steps 4 --warmup 3 --no-cpu-baseline --no-extra --no-e2e 2>>gpurun_out/r2_15_err.log > gpurun_out/r2_15_$1.json
  python -c "
import json; d=json.load(open('gpurun_out/r2_15_$1.json')); print('$1', round(d['ms_per_step'],2), '%.4g'%d['value'], round(d['roofline']['frac'],4), d['config'].get('kernel_ms'))"
}
for rep in a b; do
run particle_512_$rep librapt_b200.so particle
run particle_256_$rep librapt_b200_p256.so particle
run particle_128_$rep librapt_b200_p128.so particle
run gc_128_$rep librapt_b200.so gc
run gc_512_$rep librapt_b200_gc512.so gc
run gc_256_$rep librapt_b200_gc256.so gc
run belt_128_$rep librapt_b200.so belt
run belt_512_$rep librapt_b200_gc512.so belt
run belt_256_$rep librapt_b200_gc256.so belt
done
RAPT_B200_LIB=$PWD/rapt_b200/librapt_b200_gc512.so python bench.py --workload adaptive --steps 2 --warmup 1 --no-cpu-baseline 2>>gpurun_out/r2_15_err.log > gpurun_out/r2_15_adaptive_gc512.json; python -c "
import json; d=json.load(open('gpurun_out/r2_15_adaptive_gc512.json')); print('adaptive gc512', d['ms_per_step'], d['config']['kernel_ms'])"
python bench.py --workload adaptive --steps 2 --warmup 1 --no-cpu-baseline 2>>gpurun_out/r2_15_err.log > gpurun_out/r2_15_adaptive.json; python -c "
import json; d=json.load(open('gpurun_out/r2_15_adaptive.json')); print('adaptive default', d['ms_per_step'], d['config']['kernel_ms'])"
RAPT_B200_LIB=$PWD/rapt_b200/librapt_b200_trace.so python tools/tail_profile.py 1048576 gpurun_out/r2_15_tail_512.npz > gpurun_out/r2_15_tail_512.json 2>>gpurun_out/r2_15_err.log; python tools/tail_analyze.py gpurun_out/r2_15_tail_512.npz
python bench.py --steps 5 --warmup 3 2>>gpurun_out/r2_15_err.log > gpurun_out/r2_15_bench_n1.json; python -c "
import json; d=json.load(open('gpurun_out/r2_15_bench_n1.json')); print('bench', d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value']); print({k:(v['ms_per_step'], v['roofline']['frac']) for k,v in d['extra']['workloads'].items()})"
tail -3 gpurun_out/r2_15_err.log | cut -c1-300
