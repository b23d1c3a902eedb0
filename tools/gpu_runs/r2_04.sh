# Round 2, GPU call 4 (1 GPU): suite on the current library; north_star target on one GPU (10,485,760 protons x 10 s);
# A/B of two particle-kernel micro-variants; adaptive per-epoch breakdown.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_04_pytest.log 2>&1; tail -4 gpurun_out/r2_04_pytest.log | cut -c1-300
python bench.py --n-per-gpu 10485760 --steps 2 --warmup 2 --no-cpu-baseline 2>gpurun_out/r2_04_err.log > gpurun_out/r2_04_bench_n1_10M.json; cut -c1-250 gpurun_out/r2_04_bench_n1_10M.json
for v in "" _ctrl1 _opt2; do
  RAPT_B200_LIB=$PWD/rapt_b200/librapt_b200$v.so python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-extra 2>>gpurun_out/r2_04_err.log > gpurun_out/r2_04_bench_particle$v.json; python -c "
import json; d=json.load(open('gpurun_out/r2_04_bench_particle$v.json')); print('variant [$v]', d['ms_per_step'], d['value'], d['roofline']['frac'])"
done
for v in "" _ctrl1 _opt2; do
  RAPT_B200_LIB=$PWD/rapt_b200/librapt_b200$v.so python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-extra 2>>gpurun_out/r2_04_err.log > gpurun_out/r2_04_bench_particle${v}_b.json; python -c "
import json; d=json.load(open('gpurun_out/r2_04_bench_particle${v}_b.json')); print('variant [$v] again', d['ms_per_step'], d['value'], d['roofline']['frac'])"
done
python - > gpurun_out/r2_04_adaptive_epochs.json 2>>gpurun_out/r2_04_err.log <<'P'
import json, sys, time
import numpy as np
sys.path.insert(0, '.')
from rapt_b200 import engine, fields, synth, _lib
_lib.init(0)
out = {}
for n in (1 << 20, 1 << 16):
    ic = synth.config4_speiser(n)
    pos = np.column_stack([ic["x"], ic["y"], ic["z"]]); vel = np.column_stack([ic["vx"], ic["vy"], ic["vz"]])
    for rep in range(2):
        t = time.perf_counter()
        r = engine.adaptive_advance(fields.Parabolic(), pos, vel, ic["t0"], ic["mass"], ic["charge"], 300.0, 1.0, store_every=0,
                                    max_rows=0, arith="fast", solvertolerances=(1e-12, 1e-12), epss=0.02)
        wall = time.perf_counter() - t
    out[str(n)] = dict(wall=wall, stats=r["stats"], per_epoch=r["per_epoch"].tolist(), ok=int((r["status"] == 1).sum()),
                       nseg_hist=np.bincount(r["nseg"]).tolist(), status_hist={str(k): int(v) for k, v in zip(*np.unique(r["status"], return_counts=True))})
print(json.dumps(out))
P
python -c "
import json; d=json.load(open('gpurun_out/r2_04_adaptive_epochs.json'))
for n,v in d.items():
    print(n, v['wall'], v['stats']); pe=v['per_epoch']; print([[int(a),int(b),round(c,1),round(e,1)] for a,b,c,e in pe[:12]], '...', [[int(a),int(b),round(c,2),round(e,2)] for a,b,c,e in pe[-5:]])
"
tail -3 gpurun_out/r2_04_err.log | cut -c1-300
# ncu of the particle kernel inside an adaptive run (65,536 tracers; the first three particle-kernel launches)
cat > /tmp/adapt_small.py <<'P'
import sys, numpy as np
sys.path.insert(0, '.')
from rapt_b200 import engine, fields, synth, _lib
_lib.init(0)
n = 1 << 16
ic = synth.config4_speiser(n)
pos = np.column_stack([ic["x"], ic["y"], ic["z"]]); vel = np.column_stack([ic["vx"], ic["vy"], ic["vz"]])
r = engine.adaptive_advance(fields.Parabolic(), pos, vel, ic["t0"], ic["mass"], ic["charge"], 300.0, 1.0, store_every=0, max_rows=0,
                            arith="fast", solvertolerances=(1e-12, 1e-12), epss=0.02)
print(r["stats"])
P
ncu --set full --clock-control none --import-source on -k regex:k_particle_rkn -c 2 -o gpurun_out/r2_04_adaptive_particle python /tmp/adapt_small.py > gpurun_out/r2_04_ncu_adapt.log 2>&1
tail -2 gpurun_out/r2_04_ncu_adapt.log | cut -c1-300
