# Round 2, GPU call 2: suite on the new library (failed-row fix, diagnostics kernel, sharded ensembles, getters, long
# fixtures), the restructured bench (product multi-GPU path, reference timed in the run, extra.workloads), the reference arm,
# and the counter-parity report on the device.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/r2_02_pytest.log 2>&1; tail -25 gpurun_out/r2_02_pytest.log | cut -c1-400
python bench.py --steps 3 --warmup 3 2>gpurun_out/r2_02_bench_err.log > gpurun_out/r2_02_bench_n1.json; cut -c1-600 gpurun_out/r2_02_bench_n1.json; tail -5 gpurun_out/r2_02_bench_err.log | cut -c1-300
python bench.py --impl reference --steps 1 --warmup 1 2>>gpurun_out/r2_02_bench_err.log > gpurun_out/r2_02_bench_ref.json; cut -c1-700 gpurun_out/r2_02_bench_ref.json
python tools/count_parity_report.py --backend gpu --n 4096 --delta 0.25 > gpurun_out/r2_02_counts_4096x0.25.json 2>>gpurun_out/r2_02_bench_err.log; cut -c1-900 gpurun_out/r2_02_counts_4096x0.25.json
python tools/count_parity_report.py --backend gpu --n 2048 --delta 10 > gpurun_out/r2_02_counts_2048x10.json 2>>gpurun_out/r2_02_bench_err.log; cut -c1-900 gpurun_out/r2_02_counts_2048x10.json
python tools/count_parity_report.py --backend gpu --config 3 --n 2048 --delta 10 > gpurun_out/r2_02_counts_cfg3.json 2>>gpurun_out/r2_02_bench_err.log; cut -c1-700 gpurun_out/r2_02_counts_cfg3.json
python tools/count_parity_report.py --backend gpu --config 5 --n 2048 --delta 10 > gpurun_out/r2_02_counts_cfg5.json 2>>gpurun_out/r2_02_bench_err.log; cut -c1-700 gpurun_out/r2_02_counts_cfg5.json
tail -5 gpurun_out/r2_02_bench_err.log | cut -c1-300
