#!/bin/bash
# A/B of the two fan-plane kernels on a GPU box: parity tests, then timing of one C2 subset.
#   gpurun -- bash tools/gpu_fan_ab.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== parity (event kernel)"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -q -m gpu -k "fan" 2>&1 | tail -15
echo "== timing"
for cfg in "ev 4 8 128" "ev 4 8 256" "ev 4 8 64" "ev 4 4 128"; do
  set -- $cfg
  echo "-- impl=$1 CW=$2 NW=$3 EP=$4"
  PAX_FAN_IMPL=$1 PAX_FAN_CW=$2 PAX_FAN_NW=$3 PAX_FAN_EPOCH=$4 timeout 300 python tools/profile_step.py --compare --subsets 1 --reps 2 2>&1 | grep -v "^projector" | tail -6
done
echo "== ncu"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_ax_fan_ev -s 2 -c 1 -f -o gpurun_out/prof_ev_v1 python tools/profile_step.py --subsets 1 --reps 1 > gpurun_out/ncu_ev_v1.log 2>&1; tail -3 gpurun_out/ncu_ev_v1.log
