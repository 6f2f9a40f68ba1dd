run() { env "$@" python tools/profile_step.py --workload ${WL:-cfg5} --warm 2 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); km=d['kernels_ms']
print(' '.join('%s=%.3f'%(k,km[k]) for k in ('prior_tiles','predict_fused','assemble_A','leaf_q')), 'sum=%.2f'%sum(km.values()), repr(d['likelihood']))"; }
echo predict128; run MRA_TUNE=0
echo predict64; run MRA_TUNE=1024
export WL=cfg3
echo cfg3 predict128; run MRA_TUNE=0
echo cfg3 predict64; run MRA_TUNE=1024
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_stream.py tests/test_gpu_shard.py -m gpu -x -q 2>&1 | tail -2
