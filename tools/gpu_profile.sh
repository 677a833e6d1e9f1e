#!/bin/bash
# Evidence for profiles/: the bench (plain), then the ncu launch list of the same command, then one ncu --set full
# capture of the dominant kernel (k_ax_fan, residual mode) and of the backprojector.
#   gpurun -- bash tools/gpu_profile.sh [tag]
cd "$(dirname "$0")/.."
TAG=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary --no-e2e"
echo "== plain: $CMD"
timeout 600 $CMD > gpurun_out/plain_$TAG.json 2> gpurun_out/plain_$TAG.err && tail -c 600 gpurun_out/plain_$TAG.json
echo "== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1; tail -2 gpurun_out/ncu_launches_$TAG.log | cut -c1-300
echo "== full capture: fan projector (residual) and backprojector"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_ax_fan -s 1 -c 1 -f -o gpurun_out/prof_fan_$TAG python tools/profile_step.py --subsets 1 --reps 1 > gpurun_out/ncu_fan_$TAG.log 2>&1; tail -2 gpurun_out/ncu_fan_$TAG.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_atb -s 1 -c 1 -f -o gpurun_out/prof_atb_$TAG python tools/profile_step.py --subsets 1 --reps 2 > gpurun_out/ncu_atb_$TAG.log 2>&1; tail -2 gpurun_out/ncu_atb_$TAG.log
