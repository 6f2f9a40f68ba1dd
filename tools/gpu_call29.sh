bash tools/run_gpu_round.sh r05 k_
