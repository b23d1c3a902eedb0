N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload gc --steps 2 --warmup 2 2>gpurun_out/bench_gc_n${N}_err.log | tee gpurun_out/bench_r1_gc_n${N}.json | cut -c1-600
tail -3 gpurun_out/bench_gc_n${N}_err.log
