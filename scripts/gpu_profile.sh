#!/bin/bash
# ncu evidence for the fused tile kernel: launch list + one full capture. Run only after gpu_check.sh passed.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-graph"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:loss_strip -s 4 -c 1 -o gpurun_out/prof -f $CMD > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out; tail -3 gpurun_out/ncu_full.log
