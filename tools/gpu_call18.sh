mkdir -p gpurun_out
run() { env "$@" python tools/profile_step.py --workload cfg5 --warm 2 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); km=d['kernels_ms']
print(' '.join('%s=%.3f'%(k,km[k]) for k in ('prior_tiles','predict_fused','assemble_A','leaf_q')), 'sum=%.2f'%sum(km.values()))"; }
echo base; run A=1
echo prior_pad_24k_3ctas; run MRA_SMEM_PAD_PRIOR=24000
echo prior_pad_60k_2ctas; run MRA_SMEM_PAD_PRIOR=60000
echo predict_pad_50k_2ctas; run MRA_SMEM_PAD_PREDICT=50000
