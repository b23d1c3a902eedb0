# Round 2, GPU call 24 (1 GPU, the round's last GPU minutes): suite on the library with the history-ordered guiding-centre
# queue (sort_by_work = 2 in rapt_b200_gc_advance_dev), then config 5 and config 3 with and without it.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_24_pytest.log 2>&1; tail -3 gpurun_out/r2_24_pytest.log | cut -c1-300
for w in belt gc; do for wo in previous predicted; do
timeout 60 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --work-order $wo 2>>gpurun_out/r2_24_err.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$w $wo', d['ms_per_step'], d['value'], d['roofline']['frac']); open('gpurun_out/r2_24_ab.jsonl','a').write(json.dumps(d)+'\n')"
done; done
tail -2 gpurun_out/r2_24_err.log | cut -c1-300
