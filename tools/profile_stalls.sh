python tools/profile_step.py --workload cfg5 > gpurun_out/pp.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:k_predict_fused" -c 1 -o gpurun_out/pf -f python tools/profile_step.py --workload cfg5 > gpurun_out/ncu_pf.log 2>&1
ncu -i gpurun_out/pf.ncu-rep --page source --csv > /tmp/pf_source.csv 2>/dev/null
head -c 3000 /tmp/pf_source.csv > gpurun_out/pf_source_head.txt
python tools/top_stalls.py 60 < /tmp/pf_source.csv > gpurun_out/pf_top_stalls.txt 2>&1
ls -la gpurun_out/pf.ncu-rep /tmp/pf_source.csv; rm -f gpurun_out/pf.ncu-rep
