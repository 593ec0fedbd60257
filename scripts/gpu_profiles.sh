#!/bin/bash
# Round-2 profile captures on one B200 (every capture only after the plain command has exited 0).  Outputs in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-lbfgs --e2e-steps 1"
$CMD > gpurun_out/plain_r2.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_r2.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_steps3.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
# one evaluation of every kernel with the full set (skip the set-up and warm-up launches: capture a window late in the run)
ncu --set full --import-source on --clock-control none -k regex:"k_" -s 150 -c 24 -o gpurun_out/prof_r2_full -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full set rc=$?"
python scripts/tc_timeline.py > gpurun_out/r2_tc_timeline.txt 2>&1; echo "tc timeline rc=$?"
python scripts/graph_timeline.py c3 > gpurun_out/r2_graph_timeline.txt 2>&1; echo "graph timeline rc=$?"
python scripts/train_bench.py 12 > gpurun_out/r2_train_loop_wordnet.txt 2>&1; echo "train loop rc=$?"; tail -1 gpurun_out/r2_train_loop_wordnet.txt
