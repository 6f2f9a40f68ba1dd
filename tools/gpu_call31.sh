mkdir -p gpurun_out
MRA_HOST_TRACE=1 python tools/e2e_ab.py --workload cfg5 --reps 6 --env MRA_NO_HOST_POOL=1 --env MRA_NO_HOST_POOL=0 > gpurun_out/e2e_ab_pool.jsonl 2>gpurun_out/e2e_ab_pool.err; cat gpurun_out/e2e_ab_pool.jsonl | cut -c1-900
grep "build_lists: total\|plan_obs: total\|set_structure: total" gpurun_out/e2e_ab_pool.err | tail -12
