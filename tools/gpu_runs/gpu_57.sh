# bounce-centre kernel: lane-interleaved curves -- parity tests, then the 65 536-electron run
python -m pytest tests/test_gpu_bc.py tests/test_gpu_quad.py tests/test_gpu_gc.py -m gpu -q 2>&1 | tail -3 | cut -c1-300
timeout 600 python tools/bench_bc.py 65536 1.0 fast 2>&1 | tail -1 | tee gpurun_out/bench_bc_65536_interleaved.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('interleaved 65536', d['wall_s'], d['rhs_per_s'], d['ok_fraction'])"
