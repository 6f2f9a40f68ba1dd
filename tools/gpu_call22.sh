mkdir -p gpurun_out
MRA_HOST_TRACE=1 python tools/e2e_ab.py --workload cfg5 --reps 3 --env PYMRA_B200_TWO_PART=1 > gpurun_out/e2e_trace.jsonl 2> gpurun_out/e2e_trace.err; echo rc=$?
grep "build_lists\|plan_tree\|set_structure" gpurun_out/e2e_trace.err | tail -24
