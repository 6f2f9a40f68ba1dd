mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 --e2e-steps 5 > gpurun_out/bench_cfg5_2gpu_r05.json 2> gpurun_out/bench_cfg5_2gpu_r05.err; echo bench2_exit=$?
tail -c 200 gpurun_out/bench_cfg5_2gpu_r05.err
