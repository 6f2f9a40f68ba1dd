# Usage: bash tools/profile_stalls.sh <tag> <kernel regex> <skip> ; writes gpurun_out/<tag>_top_stalls.txt (small)
TAG=$1; KRE=$2; SKIP=${3:-0}
python tools/profile_step.py --workload cfg5 > gpurun_out/pp_$TAG.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$KRE" -s $SKIP -c 1 -o /tmp/pf_$TAG -f python tools/profile_step.py --workload cfg5 > gpurun_out/ncu_$TAG.log 2>&1
ncu -i /tmp/pf_$TAG.ncu-rep --page source --csv > /tmp/pf_$TAG.csv 2>/dev/null
python tools/top_stalls.py 45 < /tmp/pf_$TAG.csv > gpurun_out/${TAG}_top_stalls.txt 2>&1
ncu -i /tmp/pf_$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
