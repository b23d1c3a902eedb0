# Round 2, GPU call 20 (8 GPUs): speed-weighted shards (bench.py --rebalance 2) against round-robin shards, weak N = 8, same box,
# back to back, and the N = 1 line of that box.
mkdir -p gpurun_out
show() { python - "$1" <<'P'
import json, sys
d = json.load(open(sys.argv[1])); pr = d.get('per_rank') or {}
print(sys.argv[1].split('/')[-1], '%.4g' % d['value'], round(d['ms_per_step'], 1), 'frac', round(d['roofline']['frac'], 4))
print('  kernel', pr.get('kernel_ms_per_step')); print('  collect', pr.get('collect_ms_per_step')); print('  sizes', pr.get('shard_sizes'))
for c in pr.get('shard_calibration') or []: print('  calib', c)
P
}
run8() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 4 --warmup 3 --no-e2e "$@"; }
run8 --rebalance 2 2>>gpurun_out/r2_20_err.log > gpurun_out/r2_20_bench_n8_weak_rebalanced.json; show gpurun_out/r2_20_bench_n8_weak_rebalanced.json
run8 --rebalance 0 2>>gpurun_out/r2_20_err.log > gpurun_out/r2_20_bench_n8_weak_roundrobin.json; show gpurun_out/r2_20_bench_n8_weak_roundrobin.json
timeout 120 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra --no-e2e 2>>gpurun_out/r2_20_err.log > gpurun_out/r2_20_bench_n1.json; python -c "
import json; d=json.load(open('gpurun_out/r2_20_bench_n1.json')); print('n1', d['ms_per_step'], d['value'])"
grep -v "^\[e2e\|OMP_NUM\|^\*\*\*\|^$\|NCCL version" gpurun_out/r2_20_err.log | tail -8 | cut -c1-300
