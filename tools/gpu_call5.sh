mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$?
tail -4 gpurun_out/pytest_gpu.log
python tools/mle_cfg4.py > gpurun_out/mle_cfg4.jsonl 2> gpurun_out/mle_cfg4.err; echo mle_exit=$?
cat gpurun_out/mle_cfg4.jsonl
