python tools/d2h_test.py
MRA_BENCH_TRACE=1 python tools/e2e_ab.py --workload cfg5 --reps 4 --env A=1 2>&1 | tail -3
