mkdir -p gpurun_out
python bench.py --workload cfg3 --no-cpu-baseline > gpurun_out/bench_cfg3_head.json 2>/dev/null; echo b3=$?
python bench.py --workload cfg4 --steps 20 --no-cpu-baseline > gpurun_out/bench_cfg4_head.json 2>/dev/null; echo b4=$?
timeout 200 python tools/mle_cfg4.py > gpurun_out/mle_cfg4_head.jsonl 2> gpurun_out/mle_cfg4_head.err; echo mle=$?; cat gpurun_out/mle_cfg4_head.jsonl | cut -c1-300
