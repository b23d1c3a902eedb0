# guiding-centre kernel: 5 / 6 resident blocks per SM (96 / 80 registers, 20 / 24 warps)
L=$PWD/rapt_b200/librapt_b200_mb.so
for b in 5 6; do
RAPT_B200_GC_BLOCKS=$b RAPT_B200_LIB=$L python bench.py --workload gc --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>>gpurun_out/bench_err.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('gc blocks=$b', d['value'], d['ms_per_step'], d['roofline']['frac'])"
RAPT_B200_GC_BLOCKS=$b RAPT_B200_LIB=$L python bench.py --workload belt --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>>gpurun_out/bench_err.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('belt blocks=$b', d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
