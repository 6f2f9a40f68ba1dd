mkdir -p gpurun_out
run() { env "$@" python tools/profile_step.py --workload cfg5 --warm 2 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); km=d['kernels_ms']
print(' '.join('%s=%.3f'%(k,km[k]) for k in ('prior_tiles','predict_fused','fold')), 'sum=%.2f'%sum(km.values()), repr(d['likelihood']))"; }
echo 4ctas; run A=1
echo 3ctas; run MRA_SMEM_PAD_PREDICT=18000
MRA_SMEM_PAD_PREDICT=18000 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none --profile-from-start off -k regex:k_predict_fused2 -c 1 --csv --log-file gpurun_out/predict3.csv python tools/profile_step.py --workload cfg5 > /dev/null 2>&1
grep -o '"dram__bytes[^,]*","[^"]*","[^"]*"\|"lts__t_sector_hit_rate.pct","[^"]*","[^"]*"\|"gpu__time_duration.sum","[^"]*","[^"]*"' gpurun_out/predict3.csv
