#!/bin/bash
# quick per-phase timing of one or more bench configs: scripts/phase_probe.sh c3 c4
for cfg in "$@"; do
  timeout 300 python bench.py --config $cfg --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 3 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$cfg', 'ms/eval %.4f' % j['ms_per_step'], 'e2e ms %.3f' % (1e3*j['config']['global_batch']/j['e2e']['value']) if j.get('e2e') else '', ' '.join('%s=%.1f' % (k, 1e3*v['ms']) for k,v in j['phase'].items()))
"
done
