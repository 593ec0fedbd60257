#!/bin/bash
# A/B of one environment switch on the headline bench: scripts/ab_env.sh VAR v1 v2 ...   (prints ms_per_step and the phase table)
cd "$(dirname "$0")/.."
VAR="$1"; shift
for V in "$@"; do
  env "$VAR=$V" timeout 600 python bench.py --no-cpu-baseline --no-lbfgs --steps 2000 --e2e-steps 2 > gpurun_out/ab_${VAR}_$V.json 2> gpurun_out/ab_${VAR}_$V.err
  python - <<PY
import json
try:
    b=json.load(open('gpurun_out/ab_${VAR}_$V.json'))
    print('$VAR=$V', 'ms_per_step %.4f' % b['ms_per_step'], ' '.join('%s %.1f' % (k, v['ms']*1e3) for k,v in b['phase'].items()))
except Exception as e:
    print('$VAR=$V failed', e); print(open('gpurun_out/ab_${VAR}_$V.err').read()[-1500:])
PY
done
