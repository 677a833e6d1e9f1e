#!/bin/bash
# Full single-GPU check on a GPU box: the GPU test suite, smoke(), the AUTO crossover and the bench.
#   gpurun -- bash tools/gpu_full.sh [tag]
cd "$(dirname "$0")/.."
TAG=${1:-full}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
if [ -z "$SKIP_CROSSOVER" ]; then
echo "== crossover"
timeout 600 python tools/bench_crossover.py 2>&1 | tail -8
fi
echo "== bench"
timeout 1200 python bench.py $BENCH_ARGS > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$TAG.json')); print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'ax ms', d['roofline']['avg_launch_ms'], 'share', d['roofline']['share_of_step'], 'secondary', d.get('secondary'))"; tail -2 gpurun_out/bench_$TAG.err
