python -m pytest tests/test_gpu_gc.py tests/test_gpu_properties.py -m gpu -q -k "pa90 or full_size" 2>&1 | grep -E "^E|assert|Error|passed|failed" | head -60
python tools/quick_bench.py 1048576 10.0 fast 2 0 1
RAPT_B200_LIB=/root/repo/rapt_b200/librapt_b200_unr.so python tools/quick_bench.py 1048576 10.0 fast 2 0 1
