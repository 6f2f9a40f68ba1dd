mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --profile-from-start off -k regex:"k_cov_fill|k_prior_groups" --csv --log-file gpurun_out/prior_launches.csv python tools/profile_step.py --workload cfg5 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/prior_launches.csv')) if len(r)>10]
h=rows[0]; ik=h.index('Kernel Name'); im=h.index('Metric Name'); iv=h.index('Metric Value'); iid=h.index('ID')
d={}
for r in rows[1:]:
    d.setdefault((r[iid], r[ik][:28]),{})[r[im].split('.')[0][-12:]]=r[iv]
for k,v in d.items(): print(k, v)
PY
