# Round 2, GPU call 11 (8 GPUs): the weak-scaling headline with gather(sync=False) -- does the step become
# max-over-ranks(kernel) + ~1.2 ms?  Also N = 2 and 4 for the scaling profile, and the 2-rank NCCL identity test.
mkdir -p gpurun_out
show() { python - "$1" <<'P'
import json, sys
d = json.load(open(sys.argv[1])); pr = d.get('per_rank') or {}
print(sys.argv[1].split('/')[-1], '%.4g' % d['value'], round(d['ms_per_step'], 1), 'frac', round(d['roofline']['frac'], 4))
print('  kernel', pr.get('kernel_ms_per_step')); print('  collect', pr.get('collect_ms_per_step'))
P
}
python -m pytest tests/test_gpu_dist.py -q -p no:cacheprovider 2>&1 | tail -2
for n in 8 4 2 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 4 --warmup 3 --no-e2e 2>>gpurun_out/r2_11_err.log > gpurun_out/r2_11_bench_n${n}_weak_$RANDOM.json
done
for f in gpurun_out/r2_11_bench_n*_weak_*.json; do show $f; done
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra 2>>gpurun_out/r2_11_err.log > gpurun_out/r2_11_bench_n1.json; python -c "
import json; d=json.load(open('gpurun_out/r2_11_bench_n1.json')); print('n1', d['ms_per_step'], d['value'])"
grep -v "^\[e2e\|OMP_NUM\|^\*\*\*\|^$\|NCCL version" gpurun_out/r2_11_err.log | tail -5 | cut -c1-300
