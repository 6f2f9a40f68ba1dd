mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_shard.py tests/test_gpu_stream.py -m gpu -x -q > gpurun_out/pytest_gpu_2gpu.log 2>&1; echo pytest_exit=$?
tail -3 gpurun_out/pytest_gpu_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_cfg5_2gpu.json 2> gpurun_out/bench_cfg5_2gpu.err; echo bench2_exit=$?
tail -c 300 gpurun_out/bench_cfg5_2gpu.err
