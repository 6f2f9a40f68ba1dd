mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_cfg5_8gpu.json 2> gpurun_out/bench_cfg5_8gpu.err; echo bench8_exit=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_cfg5_4gpu.json 2> gpurun_out/bench_cfg5_4gpu.err; echo bench4_exit=$?
tail -c 400 gpurun_out/bench_cfg5_8gpu.err
