# Round 2, GPU call 6 (1 GPU): A/B inside one call -- work-order key (rows vs rows x steps/row predictor), nvidia-smi polling
# interval of the clock sampler (does it perturb the timed kernel?), adaptive per-epoch split after the predicate change.
mkdir -p gpurun_out
run() { # tag lib smi_ms
  RAPT_BENCH_SMI_MS=$3 RAPT_B200_LIB=$PWD/rapt_b200/$2 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra 2>>gpurun_out/r2_06_err.log > gpurun_out/r2_06_$1.json
  python -c "
import json; d=json.load(open('gpurun_out/r2_06_$1.json')); print('$1', round(d['ms_per_step'],2), '%.4g'%d['value'], round(d['roofline']['frac'],4), 'e2e %.4g'%d['e2e']['value'], d['clocks']['samples'])"
}
run rows_a librapt_b200.so 100
run key_a librapt_b200_next.so 100
run rows_b librapt_b200.so 100
run key_b librapt_b200_next.so 100
run rows_smi1000 librapt_b200.so 1000
run key_smi1000 librapt_b200_next.so 1000
run rows_smi25 librapt_b200.so 25
python - > gpurun_out/r2_06_adaptive_epochs.json 2>>gpurun_out/r2_06_err.log <<'P'
import json, sys, time
import numpy as np
sys.path.insert(0, '.')
from rapt_b200 import engine, fields, synth, _lib
_lib.init(0)
out = {}
n = 1 << 20
ic = synth.config4_speiser(n)
pos = np.column_stack([ic["x"], ic["y"], ic["z"]]); vel = np.column_stack([ic["vx"], ic["vy"], ic["vz"]])
for rep in range(2):
    t = time.perf_counter()
    r = engine.adaptive_advance(fields.Parabolic(), pos, vel, ic["t0"], ic["mass"], ic["charge"], 300.0, 1.0, store_every=0,
                                max_rows=0, arith="fast", solvertolerances=(1e-12, 1e-12), epss=0.02)
    wall = time.perf_counter() - t
out[str(n)] = dict(wall=wall, stats=r["stats"], per_epoch=r["per_epoch"].tolist(), ok=int((r["status"] == 1).sum()))
print(json.dumps(out))
P
python -c "
import json; d=json.load(open('gpurun_out/r2_06_adaptive_epochs.json'))
for n,v in d.items():
    print(n, v['wall'], v['stats']); pe=v['per_epoch']; print([[int(a),int(b),round(c,1),round(e,1)] for a,b,c,e in pe[:14]], '...', [[int(a),int(b),round(c,2),round(e,2)] for a,b,c,e in pe[-5:]])
"
tail -3 gpurun_out/r2_06_err.log | cut -c1-300
