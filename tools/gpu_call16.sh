mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$?
tail -3 gpurun_out/pytest_gpu.log
MRA_HOST_TRACE=1 python bench.py --steps 3 --warmup 3 --e2e-steps 5 --no-cpu-baseline > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo bench_exit=$?
