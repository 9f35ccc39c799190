#!/bin/bash
# ncu evidence for the fused strip kernel: launch list + one full capture. Run only after the plain run passed.
# usage: gpu_profile.sh [lib]   (lib: optional tuning variant .so)
set -u
mkdir -p gpurun_out
[ $# -gt 0 ] && export SNDE_LIB=$PWD/$1
CMD="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-graph --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:loss_strip -s 4 -c 1 -o gpurun_out/prof -f $CMD > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -4; tail -2 gpurun_out/ncu_full.log
