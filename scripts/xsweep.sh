#!/bin/bash
# sweep of the exchange's knobs on N GPUs: scripts/xsweep.sh N
cd "$(dirname "$0")/.."
N="${1:-2}"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for C in 1 2 4; do for B in 4 16 64; do
  echo "chunks=$C blocks=$B"
  NTN_EXCHANGE_CHUNKS=$C NTN_EXCHANGE_BLOCKS=$B timeout 300 $TR --master-port 29513 scripts/exchange_check.py push 2>/dev/null | grep "push:" | sed 's/.*ms\/eval/ms\/eval/'
done; done
