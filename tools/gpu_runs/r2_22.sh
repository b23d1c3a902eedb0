# Round 2, GPU call 22 (1 GPU): work order from the previous call's step counts (sort_by_work = 2, library build with the
# hint in k_particle_dt) against the predicted order, config 2, same box; GPU suite on that library.
mkdir -p gpurun_out
export RAPT_B200_LIB=$PWD/rapt_b200/librapt_b200_hint.so
for wo in previous predicted previous predicted; do
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra --no-e2e --work-order $wo 2>>gpurun_out/r2_22_err.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$wo', d['ms_per_step'], d['value'], d['roofline']['frac']); open('gpurun_out/r2_22_ab.jsonl','a').write(json.dumps(d)+'\n')"
done
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_22_pytest.log 2>&1; tail -6 gpurun_out/r2_22_pytest.log | cut -c1-300
tail -3 gpurun_out/r2_22_err.log | cut -c1-300
