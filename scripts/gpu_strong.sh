#!/bin/bash
# Strong scaling of one batch job over N GPUs: scripts/gpu_strong.sh N config [batch] [steps]
cd "$(dirname "$0")/.."
N="$1"; CFG="${2:-c4}"; B="${3:-}"; STEPS="${4:-300}"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for X in fused library; do
  OUT=gpurun_out/strong_${CFG}${B:+_b$B}_n${N}_$X
  timeout 900 $TR --master-port 29512 bench.py --gpus $N --config $CFG ${B:+--batch $B} --scaling strong --exchange $X --steps $STEPS --warmup 5 --e2e-steps 3 --no-cpu-baseline --no-lbfgs > $OUT.json 2> $OUT.err
  echo "bench $CFG strong n=$N $X rc=$?"
  python - <<PY
import json
try:
    b=json.load(open('$OUT.json'))
    print('$X', 'ms_per_step %.4f' % b['ms_per_step'],'e2e ms %.3f' % b['e2e']['ms_per_step'],'value %.4g' % b['value'],'exchange',b['runtime']['exchange'], b['runtime']['exchange_us'])
except Exception as e:
    print('bench parse failed',e); print(open('$OUT.err').read()[-2500:])
PY
done
