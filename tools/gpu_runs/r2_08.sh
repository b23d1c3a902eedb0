# Round 2, GPU call 8 (1 GPU): second A/B of the batched-probe guiding-centre kernel (two exchange sites instead of five)
mkdir -p gpurun_out
run() { # tag lib workload
  RAPT_B200_LIB=$PWD/rapt_b200/$2 python bench.py --workload $3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>>gpurun_out/r2_08_err.log > gpurun_out/r2_08_$1.json
  python -c "
import json; d=json.load(open('gpurun_out/r2_08_$1.json')); print('$1', round(d['ms_per_step'],2), '%.4g'%d['value'], round(d['roofline']['frac'],4))"
}
for rep in a b; do
run gc_defer16_$rep librapt_b200.so gc
run gc_defer8_$rep librapt_b200_defer8.so gc
run gc_nodefer_$rep librapt_b200_nodefer.so gc
done
run belt_defer16 librapt_b200.so belt
run belt_defer8 librapt_b200_defer8.so belt
run belt_nodefer librapt_b200_nodefer.so belt
ncu --set full --clock-control none --import-source on -k regex:k_gc_dopri5 -c 1 -o gpurun_out/r2_08_gc python bench.py --workload gc --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/r2_08_ncu.log 2>&1
tail -3 gpurun_out/r2_08_err.log | cut -c1-300
