mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$?
tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 --e2e-steps 2 --no-cpu-baseline > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo bench_exit=$?
python bench.py --workload cfg3 --steps 5 --e2e-steps 2 --no-cpu-baseline > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; echo bench3_exit=$?
bash tools/profile_stalls.sh r05_prior0 k_prior_groups 0
bash tools/profile_stalls.sh r05_prior6 k_prior_groups 6
bash tools/profile_stalls.sh r05_predict k_predict_fused 0
ls -la gpurun_out
