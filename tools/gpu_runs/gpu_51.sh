# guiding-centre kernel A/B: stencil as a rolled loop over the three axes (bit-identical arithmetic)
L=$PWD/rapt_b200/librapt_b200_rolled.so
for b in 4 3; do
RAPT_B200_GC_BLOCKS=$b RAPT_B200_LIB=$L python bench.py --workload gc --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>>gpurun_out/bench_err.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('gc rolled blocks=$b', d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
RAPT_B200_LIB=$L python bench.py --workload belt --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>>gpurun_out/bench_err.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('belt rolled', d['value'], d['ms_per_step'], d['roofline']['frac'])"
