mkdir -p gpurun_out
python tools/e2e_ab.py --workload cfg5 --reps 6 --env PYMRA_B200_TWO_PART=1 --env PYMRA_B200_TWO_PART=0 > gpurun_out/e2e_ab.jsonl 2> gpurun_out/e2e_ab.err; echo rc=$?
cat gpurun_out/e2e_ab.jsonl
