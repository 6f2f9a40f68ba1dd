mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$?
tail -4 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo bench_exit=$?
python bench.py --workload cfg3 --steps 5 --no-cpu-baseline > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; echo bench3_exit=$?
python bench.py --workload cfg4 --steps 10 --no-cpu-baseline > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; echo bench4_exit=$?
python tools/profile_step.py --workload cfg5 > gpurun_out/prof_plain_cfg5.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_cfg5.csv python tools/profile_step.py --workload cfg5 > gpurun_out/ncu1.log 2>&1
echo ncu1_exit=$?
bash tools/profile_stalls.sh r05_prior6 k_prior_tiles 6
bash tools/profile_stalls.sh r05_asm_leaf k_assemble_A 0
bash tools/profile_stalls.sh r05_leafq k_leaf_q 0
ls -la gpurun_out
