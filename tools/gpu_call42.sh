mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo pytest_exit=$?; tail -2 gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --no-cpu-baseline > gpurun_out/bench_cfg5_final.json 2> gpurun_out/bench_cfg5_final.err; echo bench_exit=$?
