# Round 2, GPU call 23 (1 GPU): suite on the final library (device-resident ParticleEnsemble.advance now orders by the previous
# call's step counts), and the bench line with its sub-records (incl. particle_work_order_previous); no CPU baselines (budget).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_23_pytest.log 2>&1; tail -4 gpurun_out/r2_23_pytest.log | cut -c1-300
timeout 150 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2_23_err.log > gpurun_out/r2_23_bench_n1.json; python -c "
import json; d=json.load(open('gpurun_out/r2_23_bench_n1.json')); print('bench', d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value']); print({k:(v['ms_per_step'], v['roofline']['frac']) for k,v in d['extra']['workloads'].items()})"
tail -3 gpurun_out/r2_23_err.log | cut -c1-300
