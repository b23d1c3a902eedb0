# Round 2, GPU call 9 (1 GPU): third A/B of the batched-probe guiding-centre kernel (queue counter read lazily)
mkdir -p gpurun_out
run() { # tag lib workload
  RAPT_B200_LIB=$PWD/rapt_b200/$2 python bench.py --workload $3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>>gpurun_out/r2_09_err.log > gpurun_out/r2_09_$1.json
  python -c "
import json; d=json.load(open('gpurun_out/r2_09_$1.json')); print('$1', round(d['ms_per_step'],2), '%.4g'%d['value'], round(d['roofline']['frac'],4))"
}
for rep in a b; do
run gc_defer_$rep librapt_b200.so gc
run gc_nodefer_$rep librapt_b200_nodefer.so gc
run belt_defer_$rep librapt_b200.so belt
run belt_nodefer_$rep librapt_b200_nodefer.so belt
done
ncu --set full --clock-control none --import-source on -k regex:k_gc_dopri5 -c 1 -o gpurun_out/r2_09_gc python bench.py --workload gc --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/r2_09_ncu.log 2>&1
tail -3 gpurun_out/r2_09_err.log | cut -c1-300
