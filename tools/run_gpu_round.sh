# One GPU round: parity tests, bench (ours + reference arm), ncu launch list and one full capture.
# Usage (from the repo root on the GPU box): bash tools/run_gpu_round.sh [tag] [full-capture kernel regex]
# gpurun copies back at most 64 MiB: reports are exported to CSV on the box and big .ncu-rep files dropped.
TAG=${1:-r01}
KREGEX=${2:-k_}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo pytest_exit=$?
tail -3 gpurun_out/pytest_gpu_$TAG.log
python bench.py > gpurun_out/bench_cfg5_$TAG.log 2>&1; echo bench_exit=$?
python bench.py --workload cfg3 --no-cpu-baseline > gpurun_out/bench_cfg3_$TAG.log 2>&1; echo bench3_exit=$?
python bench.py --workload cfg4 --steps 20 --no-cpu-baseline > gpurun_out/bench_cfg4_$TAG.log 2>&1; echo bench4_exit=$?
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_$TAG.log 2>&1; echo ref_exit=$?
python tools/profile_step.py --workload cfg5 > gpurun_out/prof_plain_cfg5_$TAG.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_cfg5_$TAG.csv python tools/profile_step.py --workload cfg5 > gpurun_out/ncu1_$TAG.log 2>&1
echo ncu1_exit=$?
python tools/profile_step.py --workload cfg5 > gpurun_out/prof_plain2_cfg5_$TAG.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k "regex:$KREGEX" -c 96 -o gpurun_out/prof_cfg5_$TAG -f python tools/profile_step.py --workload cfg5 > gpurun_out/ncu2_$TAG.log 2>&1
echo ncu2_exit=$?
if [ -f gpurun_out/prof_cfg5_$TAG.ncu-rep ]; then
  ncu -i gpurun_out/prof_cfg5_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_cfg5_${TAG}_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_cfg5_$TAG.ncu-rep --page details --csv > gpurun_out/prof_cfg5_${TAG}_details.csv 2>/dev/null
  ncu -i gpurun_out/prof_cfg5_$TAG.ncu-rep --page source --csv > gpurun_out/prof_cfg5_${TAG}_source.csv 2>/dev/null
  SZ=$(du -m gpurun_out/prof_cfg5_$TAG.ncu-rep | cut -f1)
  if [ "$SZ" -gt 24 ]; then rm -f gpurun_out/prof_cfg5_$TAG.ncu-rep; echo "ncu-rep dropped ($SZ MiB)"; fi
fi
for f in gpurun_out/*; do SZ=$(du -m "$f" | cut -f1); if [ "$SZ" -gt 20 ]; then echo "dropping $f ($SZ MiB)"; rm -f "$f"; fi; done
du -sh gpurun_out; nproc; free -g | head -2
