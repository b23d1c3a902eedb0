# bounce-centre: work ordering by field-line length -- tests, then 65536 / 262144 tracers with and without it
timeout 900 python -m pytest tests/test_gpu_bc.py tests/test_gpu_quad.py -m gpu -q 2>&1 | tail -3
for srt in 0 1; do
RAPT_B200_BC_SORT=$srt timeout 600 python tools/bench_bc.py 65536 1.0 fast 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('sort=$srt 65536', d['wall_s'], d['rhs_per_s'], d['ok_fraction'])"
done
RAPT_B200_BC_SORT=1 timeout 900 python tools/bench_bc.py 262144 1.0 fast 2>&1 | tail -1 | tee gpurun_out/bench_bc_262144.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('sort=1 262144', d['wall_s'], d['rhs_per_s'], d['ok_fraction'])"
