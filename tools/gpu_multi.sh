N=$1
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 2>gpurun_out/bench_n${N}_err.log | tee gpurun_out/bench_r1_n${N}.json | cut -c1-900
tail -3 gpurun_out/bench_n${N}_err.log
