mkdir -p gpurun_out
MRA_HOST_TRACE=1 python tools/e2e_ab.py --workload cfg5 --reps 6 --env A=1 > gpurun_out/e2e_final.jsonl 2> gpurun_out/e2e_final.err; cat gpurun_out/e2e_final.jsonl | cut -c1-700
grep "build_lists: \|plan_tree: total" gpurun_out/e2e_final.err | tail -6
timeout 600 python -m pytest tests/test_gpu_stream.py tests/test_gpu_shard.py -m gpu -x -q 2>&1 | tail -2
