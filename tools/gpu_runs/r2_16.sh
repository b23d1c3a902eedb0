# Round 2, GPU call 16 (1 GPU): the final library (512-thread blocks for launches that fill the GPU, 128-thread blocks for small
# ones) -- suite, the default bench line, adaptive, 10 M protons on one GPU, ncu + launch list of the bench command.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_16_pytest.log 2>&1; tail -4 gpurun_out/r2_16_pytest.log | cut -c1-300
python bench.py --steps 5 --warmup 3 2>gpurun_out/r2_16_err.log > gpurun_out/r2_16_bench_n1.json; python -c "
import json; d=json.load(open('gpurun_out/r2_16_bench_n1.json')); print('bench', d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value']); print({k:(v['ms_per_step'], v['roofline']['frac'], v.get('kernel_ms')) for k,v in d['extra']['workloads'].items()})"
python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/r2_16_err.log > gpurun_out/r2_16_bench_ref.json; cut -c1-200 gpurun_out/r2_16_bench_ref.json
python bench.py --n-per-gpu 10485760 --steps 2 --warmup 2 --no-cpu-baseline --no-extra 2>>gpurun_out/r2_16_err.log > gpurun_out/r2_16_bench_n1_10M.json; python -c "
import json; d=json.load(open('gpurun_out/r2_16_bench_n1_10M.json')); print('10M', d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_16_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/r2_16_launches_stdout.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_particle_rkn -c 1 -o gpurun_out/r2_16_particle python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/r2_16_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_gc_dopri5 -c 1 -o gpurun_out/r2_16_gc python bench.py --workload gc --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/r2_16_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_gc_dopri5 -c 1 -o gpurun_out/r2_16_belt python bench.py --workload belt --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/r2_16_ncu3.log 2>&1
tail -3 gpurun_out/r2_16_err.log | cut -c1-300
