# guiding-centre kernel A/B: cooperative probe only when at most 4 (m4) / 8 (m8) lanes of the warp start a row
for v in _m4 _m8; do
RAPT_B200_LIB=$PWD/rapt_b200/librapt_b200$v.so python bench.py --workload gc --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>>gpurun_out/bench_err.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('gc lib$v', d['value'], d['ms_per_step'], d['roofline']['frac'])"
RAPT_B200_LIB=$PWD/rapt_b200/librapt_b200$v.so python bench.py --workload belt --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>>gpurun_out/bench_err.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('belt lib$v', d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
