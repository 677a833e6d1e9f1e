#!/bin/bash
# Multi-GPU check on an N-GPU box: sharded-vs-single tests, then the bench at N ranks.
#   gpurun --gpus N -- bash tools/gpu_multi.sh N [tag]
cd "$(dirname "$0")/.."
N=${1:-2}; TAG=${2:-n$N}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c
if [ -z "$SKIP_TESTS" ]; then
echo "== pytest tests/test_multigpu.py"
timeout 900 python -m pytest tests/test_multigpu.py -q -m gpu -s 2>&1 | tail -15
fi
echo "== bench N=$N"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N $BENCH_ARGS > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = [json.loads(l) for l in open('gpurun_out/bench_$TAG.json') if l.startswith('{')][-1]
    print('ms_per_step', d['ms_per_step'], 'phases', d.get('subset_step_ms'), 'sharding', (d.get('sharding') or {}).get('grid'))
    print('parity', json.dumps(d.get('parity_vs_1gpu')))
except Exception as e:
    print('bench failed', e)
PY
tail -3 gpurun_out/bench_$TAG.err
