mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$?
tail -5 gpurun_out/pytest_gpu.log
MRA_RUN_SLOW=1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k cfg3 > gpurun_out/pytest_gpu_slow.log 2>&1; echo pytest_slow_exit=$?
tail -3 gpurun_out/pytest_gpu_slow.log
python tools/profile_step.py --workload cfg5 > gpurun_out/prof_plain_cfg5.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_cfg5.csv python tools/profile_step.py --workload cfg5 > gpurun_out/ncu1.log 2>&1
echo ncu1_exit=$?
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k "regex:k_" -c 90 -o gpurun_out/prof_cfg5 -f python tools/profile_step.py --workload cfg5 > gpurun_out/ncu2.log 2>&1
echo ncu2_exit=$?
if [ -f gpurun_out/prof_cfg5.ncu-rep ]; then
  ncu -i gpurun_out/prof_cfg5.ncu-rep --page raw --csv > gpurun_out/prof_cfg5_raw.csv 2>/dev/null
  SZ=$(du -m gpurun_out/prof_cfg5.ncu-rep | cut -f1)
  if [ "$SZ" -gt 40 ]; then rm -f gpurun_out/prof_cfg5.ncu-rep; echo "ncu-rep dropped ($SZ MiB)"; fi
fi
du -sh gpurun_out
