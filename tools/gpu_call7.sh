mkdir -p gpurun_out
nvidia-smi -L | head -3
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_2gpu.log 2>&1; echo pytest_exit=$?
tail -4 gpurun_out/pytest_gpu_2gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5_1gpu.json 2> gpurun_out/bench_cfg5_1gpu.err; echo bench1_exit=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_cfg5_2gpu.json 2> gpurun_out/bench_cfg5_2gpu.err; echo bench2_exit=$?
tail -c 600 gpurun_out/bench_cfg5_2gpu.err
