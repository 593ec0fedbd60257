#!/bin/bash
# One GPU cycle: parity tests, then a short bench.  Usage: scripts/gpu_cycle.sh [pytest -k expression]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
K="${1:-}"
if [ -n "$K" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "$K" > gpurun_out/pytest.log 2>&1
else
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1
fi
echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -15 gpurun_out/pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"
python - <<'PY'
import json
try:
    b=json.load(open('gpurun_out/bench.json'))
    print('ms_per_step',b['ms_per_step'],'e2e ms',b['e2e']['ms_per_step'],'cost',b['cost'],'nact',b['n_active'])
    for k,v in b['phase'].items(): print(f"  {k:14s} {v['ms']*1e3:8.1f} us")
except Exception as e:
    print('bench parse failed',e); print(open('gpurun_out/bench.err').read()[-2000:])
PY
