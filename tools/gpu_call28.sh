python tools/predict_tail.py 2>&1 | grep "predict_wall" | tail -2
python tools/e2e_ab.py --workload cfg5 --reps 5 --env A=1 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
