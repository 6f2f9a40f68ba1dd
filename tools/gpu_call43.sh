mkdir -p gpurun_out
python tools/profile_step.py --workload cfg5 > gpurun_out/pp_tail.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --profile-from-start off -k "regex:k_fold|k_predict_fused2" -c 2 -o /tmp/prof_tail -f python tools/profile_step.py --workload cfg5 > gpurun_out/ncu_tail.log 2>&1
ncu -i /tmp/prof_tail.ncu-rep --page raw --csv > gpurun_out/prof_cfg5_r05f_tail_raw.csv 2>/dev/null
wc -c gpurun_out/prof_cfg5_r05f_tail_raw.csv
