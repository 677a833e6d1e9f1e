#!/bin/bash
# The fan-plane kernel on a GPU box: parity tests, best-of-6 time of one C2 subset (residual mode) for the launch
# shapes the library is built with, and one ncu --set full capture.    gpurun -- bash tools/gpu_fan_ab.sh [tag]
cd "$(dirname "$0")/.."
TAG=${1:-fan}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== parity"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -q -m gpu -k "fan" 2>&1 | tail -5
echo "== timing (ms per 20-angle C2 subset, three best of six)"
for cfg in "4 160" "8 160" "4 128" "4 192"; do
  set -- $cfg
  echo "-- warps per CTA $1, epoch $2"
  PAX_FAN_NW=$1 PAX_FAN_EPOCH=$2 timeout 300 python tools/profile_step.py --subsets 1 --reps 6 2>&1 | grep -E "rep [0-9]" | awk '{print $6}' | sort -n | head -3 | tr '\n' ' '; echo
done
echo "== ncu"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_ax_fan -s 1 -c 1 -f -o gpurun_out/prof_$TAG python tools/profile_step.py --subsets 1 --reps 1 > gpurun_out/ncu_$TAG.log 2>&1; tail -2 gpurun_out/ncu_$TAG.log
