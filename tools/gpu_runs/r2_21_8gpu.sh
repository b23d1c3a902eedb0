# Round 2, GPU call 21 (8 GPUs): the north_star strong run (10,485,760 protons over 8 GPUs) with time-weighted shards
# (bench.py default from this call on: --rebalance 2).
mkdir -p gpurun_out
timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 3 --scaling strong --n-per-gpu 10485760 --no-e2e 2>gpurun_out/r2_21_err.log > gpurun_out/r2_21_bench_n8_strong_10M.json
python - <<'P'
import json
d = json.load(open('gpurun_out/r2_21_bench_n8_strong_10M.json')); pr = d.get('per_rank') or {}
print('%.4g' % d['value'], round(d['ms_per_step'], 1), 'frac', round(d['roofline']['frac'], 4), d['config']['shards'][:40])
print('  kernel', pr.get('kernel_ms_per_step')); print('  collect', pr.get('collect_ms_per_step')); print('  sizes', pr.get('shard_sizes'))
for c in pr.get('shard_calibration') or []: print('  calib', c)
P
grep -v "^\[e2e\|OMP_NUM\|^\*\*\*\|^$\|NCCL version" gpurun_out/r2_21_err.log | tail -8 | cut -c1-300
