#!/bin/bash
# A/B of the fan-plane kernel variants on a GPU box: parity tests, timing of one C2 subset, one ncu capture.
#   gpurun -- bash tools/gpu_fan_ab.sh [tag]
cd "$(dirname "$0")/.."
TAG=${1:-ev}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== parity (event kernel)"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -q -m gpu -k "fan" 2>&1 | tail -5
echo "== timing"
for cfg in "ev 8 4 160 4" "ev 8 8 160 2" "ev 8 16 160 1"; do
  set -- $cfg
  echo "-- impl=$1 CW=$2 NW=$3 EP=$4 MINB=$5"
  PAX_FAN_IMPL=$1 PAX_FAN_CW=$2 PAX_FAN_NW=$3 PAX_FAN_EPOCH=$4 PAX_FAN_MINB=$5 timeout 300 python tools/profile_step.py --subsets 1 --reps 6 2>&1 | grep -E "rep [0-9]" | awk '{print $6}' | sort -n | head -3 | tr '\n' ' '; echo
done
echo "== ncu"
PAX_FAN_NW=16 PAX_FAN_MINB=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_ax_fan -s 1 -c 1 -f -o gpurun_out/prof_$TAG python tools/profile_step.py --subsets 1 --reps 1 > gpurun_out/ncu_$TAG.log 2>&1; tail -2 gpurun_out/ncu_$TAG.log
exit 0
echo "== bench"
timeout 900 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$TAG.json')); print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'ax ms', d['roofline']['avg_launch_ms'], 'share', d['roofline']['share_of_step'])"; tail -2 gpurun_out/bench_$TAG.err
