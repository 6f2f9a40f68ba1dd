mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_features.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$?
tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 --e2e-steps 2 --no-cpu-baseline > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo bench_exit=$?
python bench.py --workload cfg3 --steps 5 --e2e-steps 2 --no-cpu-baseline > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; echo bench3_exit=$?
python tools/parity_table.py > gpurun_out/parity_table.md 2> gpurun_out/parity_table.err; echo table_exit=$?
