# Round 2, GPU call 14 (1 GPU): is the tail of the particle kernel unfair warp scheduling?  One 512-thread block per SM with a
# block barrier every iteration (ls1) / every 4th (ls4) / none (b512) against the default 4 x 128 threads, in one call; the
# per-tracer time profile of the lockstep build.
mkdir -p gpurun_out
run() { # tag lib
  RAPT_B200_LIB=$PWD/rapt_b200/$2 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra --no-e2e 2>>gpurun_out/r2_14_err.log > gpurun_out/r2_14_$1.json
  python -c "
import json; d=json.load(open('gpurun_out/r2_14_$1.json')); print('$1', round(d['ms_per_step'],2), '%.4g'%d['value'], round(d['roofline']['frac'],4))"
}
for rep in a b; do
run default_$rep librapt_b200.so
run ls1_$rep librapt_b200_ls1.so
run ls4_$rep librapt_b200_ls4.so
run b512_$rep librapt_b200_b512.so
done
RAPT_B200_LIB=$PWD/rapt_b200/librapt_b200_ls1trace.so python tools/tail_profile.py 1048576 gpurun_out/r2_14_tail_ls1.npz > gpurun_out/r2_14_tail_ls1.json 2>>gpurun_out/r2_14_err.log; python tools/tail_analyze.py gpurun_out/r2_14_tail_ls1.npz
RAPT_B200_LIB=$PWD/rapt_b200/librapt_b200_ls1.so python -m pytest tests/test_gpu_particle.py tests/test_gpu_adaptive.py tests/test_gpu_properties.py -q -p no:cacheprovider 2>&1 | tail -2
tail -3 gpurun_out/r2_14_err.log | cut -c1-300
