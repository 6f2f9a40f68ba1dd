mkdir -p gpurun_out
python bench.py --workload cfg3 --steps 5 --e2e-steps 3 --no-cpu-baseline > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; echo bench3_exit=$?
python bench.py --workload cfg4 --steps 10 --e2e-steps 3 --no-cpu-baseline > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; echo bench4_exit=$?
python bench.py --steps 5 --warmup 3 --e2e-steps 5 --no-cpu-baseline > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo bench_exit=$?
