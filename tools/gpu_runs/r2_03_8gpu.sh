# Round 2, GPU call 3 (8 GPUs of one box): the product multi-GPU path.  2-rank NCCL identity test, the weak-scaling headline
# at N = 8 with the per-rank kernel / collection split, and the north_star target: 10,485,760 protons x 10 s split over 8 GPUs.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.max.sm,power.limit --format=csv,noheader | head -8
python -m pytest tests/test_gpu_dist.py -q -p no:cacheprovider 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 8 --steps 3 --warmup 3 2>gpurun_out/r2_03_err.log > gpurun_out/r2_03_bench_n8_weak.json; cut -c1-300 gpurun_out/r2_03_bench_n8_weak.json
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_03_bench_n8_weak.json')); print(json.dumps(d.get('per_rank'))[:1500]); print(d['value'], d['ms_per_step'], d['roofline']['frac'])
P
$TR bench.py --gpus 8 --steps 3 --warmup 3 --scaling strong --n-per-gpu 10485760 --no-e2e 2>>gpurun_out/r2_03_err.log > gpurun_out/r2_03_bench_n8_strong_10M.json; cut -c1-300 gpurun_out/r2_03_bench_n8_strong_10M.json
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_03_bench_n8_strong_10M.json')); print(json.dumps(d.get('per_rank'))[:1500]); print(d['value'], d['ms_per_step'], d['roofline']['frac'])
P
tail -5 gpurun_out/r2_03_err.log | cut -c1-300
