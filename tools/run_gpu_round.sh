# One GPU round: parity tests, bench (ours + reference arm), ncu launch list and one full capture.
# Usage (from the repo root on the GPU box): bash tools/run_gpu_round.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo pytest_exit=$?
python bench.py > gpurun_out/bench_cfg5_$TAG.log 2>&1; echo bench_exit=$?
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_$TAG.log 2>&1; echo ref_exit=$?
python tools/profile_step.py --workload cfg5 > gpurun_out/prof_plain_cfg5_$TAG.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_cfg5_$TAG.csv python tools/profile_step.py --workload cfg5 > gpurun_out/ncu1_$TAG.log 2>&1
echo ncu1_exit=$?
python tools/profile_step.py --workload cfg3 > gpurun_out/prof_plain_cfg3_$TAG.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -o gpurun_out/prof_cfg3_$TAG -f python tools/profile_step.py --workload cfg3 > gpurun_out/ncu2_$TAG.log 2>&1
echo ncu2_exit=$?
nproc; free -g | head -2
