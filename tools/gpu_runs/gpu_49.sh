# guiding-centre kernel A/B: v1 = loop restructure only (plain probe), v2 = cooperative probe as a noinline function, v3 = inlined (gpu_48)
for v in _v1 _v2 ""; do
RAPT_B200_LIB=$PWD/rapt_b200/librapt_b200$v.so python bench.py --workload gc --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>>gpurun_out/bench_err.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('gc lib$v coop=1', d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
RAPT_B200_GC_COOP=0 RAPT_B200_LIB=$PWD/rapt_b200/librapt_b200_v1.so python bench.py --workload gc --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>>gpurun_out/bench_err.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('gc lib_v1 coop=0', d['value'], d['ms_per_step'], d['roofline']['frac'])"
