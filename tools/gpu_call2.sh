mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$?
tail -5 gpurun_out/pytest_gpu.log
MRA_RUN_SLOW=1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k cfg3 > gpurun_out/pytest_gpu_slow.log 2>&1; echo pytest_slow_exit=$?
tail -3 gpurun_out/pytest_gpu_slow.log
python tools/parity_table.py > gpurun_out/r04_parity_table.md 2> gpurun_out/parity_table.err; echo table_exit=$?
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo bench_exit=$?
MRA_PRIOR_GROUPS=0 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5_nogroups.json 2> gpurun_out/bench_cfg5_nogroups.err; echo bench_ng_exit=$?
python bench.py --workload cfg3 --steps 5 --no-cpu-baseline > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; echo bench3_exit=$?
python bench.py --workload cfg4 --steps 10 --no-cpu-baseline > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; echo bench4_exit=$?
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo ref_exit=$?
MRA_CHOL_MMA=0 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5_cholscalar.json 2> gpurun_out/bench_cfg5_cholscalar.err; echo bench_cs_exit=$?
