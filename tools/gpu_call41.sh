run() { env "$@" python tools/profile_step.py --workload ${WL:-cfg5} --warm 2 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); km=d['kernels_ms']
print(' '.join('%s=%.3f'%(k,km[k]) for k in ('prior_tiles','predict_fused','assemble_A','leaf_q','leaf_ut')), 'sum=%.2f'%sum(km.values()), repr(d['likelihood']))"; }
run MRA_TUNE=0
WL=cfg3 run MRA_TUNE=0
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -2
