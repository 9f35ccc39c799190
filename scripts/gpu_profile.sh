#!/bin/bash
# ncu evidence for the fused strip kernel: launch list + one full capture per workload. Run only after the plain run passed.
# usage: gpu_profile.sh tag workload [workload ...]      (files land in gpurun_out/<tag>_<workload>_*)
set -u
mkdir -p gpurun_out
tag=$1; shift
for wl in "$@"; do
  CMD="python bench.py --workload $wl --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-train"
  $CMD > gpurun_out/${tag}_${wl}_plain.log 2>&1 || { echo "plain run failed ($wl)"; tail -5 gpurun_out/${tag}_${wl}_plain.log; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${tag}_${wl}_launches.csv $CMD > /dev/null 2>&1
  ncu --set full --clock-control none --import-source on -k regex:loss_strip -s 4 -c 1 -o gpurun_out/${tag}_${wl}_prof -f $CMD > gpurun_out/${tag}_${wl}_ncu_full.log 2>&1
  ncu -i gpurun_out/${tag}_${wl}_prof.ncu-rep --page details > gpurun_out/${tag}_${wl}_details.txt 2>&1
  ncu -i gpurun_out/${tag}_${wl}_prof.ncu-rep --page raw --csv > gpurun_out/${tag}_${wl}_raw.csv 2>&1
  [ "$wl" != "config2" ] && rm -f gpurun_out/${tag}_${wl}_prof.ncu-rep      # keep the scratch area small: only the headline report travels back
done
ls -la gpurun_out | grep ${tag}_ | head -30
