mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stream.py tests/test_gpu_parity.py tests/test_gpu_features.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$?
tail -15 gpurun_out/pytest_gpu.log
MRA_HOST_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --e2e-steps 5 --no-cpu-baseline > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo bench_exit=$?
PYMRA_B200_TWO_PART=0 timeout 600 python bench.py --steps 3 --warmup 3 --e2e-steps 5 --no-cpu-baseline > gpurun_out/bench_cfg5_onepart.json 2> gpurun_out/bench_cfg5_onepart.err; echo bench1_exit=$?
