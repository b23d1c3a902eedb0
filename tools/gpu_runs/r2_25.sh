# Round 2, GPU call 25 (1 GPU, last 1.8 GPU-minutes): suite on the library with the a-priori guiding-centre work-order key
# (k_key_gc), then the cold calls of configs 3 and 5 (call 24, member order: 212.2 / 383.5 ms; previous-call order: 194.7 / 356.5 ms).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_25_pytest.log 2>&1; tail -2 gpurun_out/r2_25_pytest.log | cut -c1-200
for w in gc belt; do
timeout 40 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --work-order predicted 2>>gpurun_out/r2_25_err.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$w cold', d['ms_per_step'], d['value'], d['roofline']['frac']); open('gpurun_out/r2_25_ab.jsonl','a').write(json.dumps(d)+'\n')"
done
