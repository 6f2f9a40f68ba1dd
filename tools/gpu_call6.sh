mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_features.py tests/test_gpu_stream.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$?
tail -4 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo bench_exit=$?
MRA_LEAF_WIDE=0 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5_narrow.json 2> gpurun_out/bench_cfg5_narrow.err; echo bench_n_exit=$?
python bench.py --workload cfg3 --steps 5 --no-cpu-baseline > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; echo bench3_exit=$?
python tools/mle_cfg4.py > gpurun_out/mle_cfg4.jsonl 2> gpurun_out/mle_cfg4.err; echo mle_exit=$?
cat gpurun_out/mle_cfg4.jsonl
