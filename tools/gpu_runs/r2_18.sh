# Round 2, GPU call 18 (1 GPU): suite on the library with the periodic shard table in rapt_b200_unshard_dev (ShardPlan).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_18_pytest.log 2>&1; tail -15 gpurun_out/r2_18_pytest.log | cut -c1-300
