# Round 2, GPU call 10 (8 GPUs of one box): the product multi-GPU path on the final library -- weak-scaling headline with
# the collection split, north_star strong run, and the two full-size guiding-centre configurations.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
show() { python - "$1" <<'P'
import json, sys
d = json.load(open(sys.argv[1])); pr = d.get('per_rank') or {}
print(sys.argv[1].split('/')[-1], '%.4g' % d['value'], round(d['ms_per_step'], 1), 'frac', round(d['roofline']['frac'], 4), 'e2e', d.get('e2e', {}).get('value'))
print('  kernel', pr.get('kernel_ms_per_step')); print('  collect', pr.get('collect_ms_per_step')); print('  split', pr.get('collect_split_ms_after_barrier'))
P
}
$TR bench.py --gpus 8 --steps 4 --warmup 3 2>gpurun_out/r2_10_err.log > gpurun_out/r2_10_bench_n8_weak.json; show gpurun_out/r2_10_bench_n8_weak.json
$TR bench.py --gpus 8 --steps 3 --warmup 3 --scaling strong --n-per-gpu 10485760 --no-e2e 2>>gpurun_out/r2_10_err.log > gpurun_out/r2_10_bench_n8_strong_10M.json; show gpurun_out/r2_10_bench_n8_strong_10M.json
$TR bench.py --gpus 8 --workload belt --n-per-gpu 12500000 --steps 2 --warmup 1 --no-e2e 2>>gpurun_out/r2_10_err.log > gpurun_out/r2_10_bench_config5_100M_8gpu.json; show gpurun_out/r2_10_bench_config5_100M_8gpu.json
$TR bench.py --gpus 8 --workload gc --scaling strong --n-per-gpu 10485760 --delta 100 --steps 2 --warmup 1 --no-e2e 2>>gpurun_out/r2_10_err.log > gpurun_out/r2_10_bench_config3_10M_100s_8gpu.json; show gpurun_out/r2_10_bench_config3_10M_100s_8gpu.json
$TR bench.py --gpus 8 --steps 3 --warmup 2 --no-e2e 2>>gpurun_out/r2_10_err.log > gpurun_out/r2_10_bench_n8_weak_b.json; show gpurun_out/r2_10_bench_n8_weak_b.json
grep -v "^\[e2e\|OMP_NUM\|^\*\*\*\|^$\|NCCL version" gpurun_out/r2_10_err.log | tail -5 | cut -c1-300
