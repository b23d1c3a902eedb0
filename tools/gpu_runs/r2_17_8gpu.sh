# Round 2, GPU call 17 (8 GPUs): the final library through the product multi-GPU path -- weak N = 8, 4, 2, 1 back to back and
# the north_star strong run (10,485,760 protons over 8 GPUs).
mkdir -p gpurun_out
show() { python - "$1" <<'P'
import json, sys
d = json.load(open(sys.argv[1])); pr = d.get('per_rank') or {}
print(sys.argv[1].split('/')[-1], '%.4g' % d['value'], round(d['ms_per_step'], 1), 'frac', round(d['roofline']['frac'], 4))
print('  kernel', pr.get('kernel_ms_per_step')); print('  collect', pr.get('collect_ms_per_step'))
P
}
for n in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 4 --warmup 3 2>>gpurun_out/r2_17_err.log > gpurun_out/r2_17_bench_n${n}_weak.json; show gpurun_out/r2_17_bench_n${n}_weak.json
done
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra 2>>gpurun_out/r2_17_err.log > gpurun_out/r2_17_bench_n1.json; python -c "
import json; d=json.load(open('gpurun_out/r2_17_bench_n1.json')); print('n1', d['ms_per_step'], d['value'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 3 --scaling strong --n-per-gpu 10485760 --no-e2e 2>>gpurun_out/r2_17_err.log > gpurun_out/r2_17_bench_n8_strong_10M.json; show gpurun_out/r2_17_bench_n8_strong_10M.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 4 --warmup 3 --no-e2e 2>>gpurun_out/r2_17_err.log > gpurun_out/r2_17_bench_n8_weak_b.json; show gpurun_out/r2_17_bench_n8_weak_b.json
grep -v "^\[e2e\|OMP_NUM\|^\*\*\*\|^$\|NCCL version" gpurun_out/r2_17_err.log | tail -5 | cut -c1-300
