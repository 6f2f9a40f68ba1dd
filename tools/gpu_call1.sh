mkdir -p gpurun_out
nproc; free -g | head -2
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$?
tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo bench_exit=$?
python bench.py --workload cfg3 --steps 5 --no-cpu-baseline > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; echo bench3_exit=$?
python tools/parity_fullsize.py cfg4 cfg4_exp g700_r16_m7 g700_r16_m7_exp > gpurun_out/parity_a.jsonl 2> gpurun_out/parity_a.err; echo pa_exit=$?
python tools/parity_fullsize.py cfg3 cfg3_exp > gpurun_out/parity_b.jsonl 2> gpurun_out/parity_b.err; echo pb_exit=$?
python tools/parity_fullsize.py cfg5_exp > gpurun_out/parity_c.jsonl 2> gpurun_out/parity_c.err; echo pc_exit=$?
python tools/parity_fullsize.py cfg5 > gpurun_out/parity_d.jsonl 2> gpurun_out/parity_d.err; echo pd_exit=$?
