echo "--- default (128 thr, MINB 4)"; python tools/quick_bench.py 1048576 10 fast 2 2>&1 | tail -1
for v in a b c; do echo "--- variant $v"; RAPT_B200_LIB=$PWD/rapt_b200/librapt_b200_$v.so python tools/quick_bench.py 1048576 10 fast 2 2>&1 | tail -1; done
