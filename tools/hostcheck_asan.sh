#!/bin/bash
# The advance kernels' source under AddressSanitizer / UBSan on the CPU: builds tests/hostcheck/_asan/*.so and runs
# tests/test_kernel_host.py against them (out-of-bounds row writes, work-list indexing, uninitialised reads ...).
set -e
cd "$(dirname "$0")/.."
make -C tests/hostcheck -s asan
cat > /tmp/_hostcheck_asan.py <<'PY'
import sys
sys.path[:0] = [".", "oracle", "tests"]
import hostkernel
hostkernel._DIR = "tests/hostcheck/_asan"
import pytest
sys.exit(pytest.main(["tests/test_kernel_host.py", "-x", "-q", "-s", "-p", "no:cacheprovider"]))
PY
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:halt_on_error=1:verify_asan_link_order=0 \
    UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 python /tmp/_hostcheck_asan.py
