# Round 2, GPU call 5 (1 GPU): suite + bench on the library with the micro-variants adopted (controller without division,
# series r^-5), the division-free adiabaticity predicates and the log/exp HINIT root; adaptive line; counter parity again.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_05_pytest.log 2>&1; tail -6 gpurun_out/r2_05_pytest.log | cut -c1-300
python bench.py --steps 4 --warmup 3 2>gpurun_out/r2_05_err.log > gpurun_out/r2_05_bench_n1.json
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_05_bench_n1.json'))
print('particle', d['ms_per_step'], d['value'], d['roofline']['frac'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'])
for k,v in d['extra']['workloads'].items(): print(k, v['ms_per_step'], v['value'], v['roofline']['frac'], v.get('kernel_ms'), v.get('epochs'))
P
python bench.py --workload adaptive --steps 2 --warmup 1 --no-cpu-baseline 2>>gpurun_out/r2_05_err.log > gpurun_out/r2_05_bench_adaptive.json; cut -c1-1500 gpurun_out/r2_05_bench_adaptive.json
python tools/count_parity_report.py --backend gpu --n 2048 --delta 10 > gpurun_out/r2_05_counts_2048x10.json 2>>gpurun_out/r2_05_err.log; cut -c1-1200 gpurun_out/r2_05_counts_2048x10.json
python tools/count_parity_report.py --backend gpu --config 3 --n 2048 --delta 10 > gpurun_out/r2_05_counts_cfg3.json 2>>gpurun_out/r2_05_err.log; cut -c1-700 gpurun_out/r2_05_counts_cfg3.json
ncu --set full --clock-control none --import-source on -k regex:k_particle_rkn -c 1 -o gpurun_out/r2_05_particle python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/r2_05_ncu.log 2>&1
tail -3 gpurun_out/r2_05_err.log | cut -c1-300
