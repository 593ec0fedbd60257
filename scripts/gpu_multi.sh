#!/bin/bash
# Multi-GPU cycle: exchange check + timeline + bench (fused vs library exchange) on N GPUs.  Usage: scripts/gpu_multi.sh N [config] [scaling]
cd "$(dirname "$0")/.."
N="${1:-2}"; CFG="${2:-c3}"; SC="${3:-weak}"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 scripts/exchange_check.py > gpurun_out/xcheck_n$N.txt 2> gpurun_out/xcheck_n$N.err
echo "xcheck rc=$?"; grep -E "push:|p2p:|multimem:" gpurun_out/xcheck_n$N.txt; grep -iE "error|Traceback" gpurun_out/xcheck_n$N.err | head -5
timeout 300 $TR --master-port 29514 scripts/xtimeline.py $CFG 2>/dev/null | grep -E "world|word_pull|exchange_|finalize|entity_pull" | tee gpurun_out/xtimeline_${CFG}_n$N.txt
for X in fused library; do
  timeout 600 $TR --master-port 29512 bench.py --gpus $N --config $CFG --scaling $SC --exchange $X --steps 2000 --no-cpu-baseline \
     > gpurun_out/bench_${CFG}_${SC}_n${N}_$X.json 2> gpurun_out/bench_${CFG}_${SC}_n${N}_$X.err
  echo "bench $X rc=$?"
  python - <<PY
import json
try:
    b=json.load(open('gpurun_out/bench_${CFG}_${SC}_n${N}_$X.json'))
    print('$X', 'ms_per_step %.4f' % b['ms_per_step'],'e2e ms %.3f' % b['e2e']['ms_per_step'],'value %.4g' % b['value'],'exchange',b['runtime']['exchange'], b['runtime']['exchange_us'])
except Exception as e:
    print('bench parse failed',e); print(open('gpurun_out/bench_${CFG}_${SC}_n${N}_$X.err').read()[-3000:])
PY
done
