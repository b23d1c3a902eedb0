N=$1; W=$2; NP=$3; D=$4
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --workload $W --n-per-gpu $NP --delta $D --steps 2 --warmup 1 --no-e2e --cpu-sample 512 2>gpurun_out/bench_${W}_big_err.log | tee gpurun_out/bench_r1_${W}_n${N}_big.json | cut -c1-700
tail -3 gpurun_out/bench_${W}_big_err.log
