#!/bin/bash
# diagnostic: the SURVEY 8(d) white-noise generator vs band-limited frames and depth (coalesced gathers) on the three shapes
mkdir -p gpurun_out; : > gpurun_out/smooth_inputs.txt
for wl in config2 nyu hires; do
  for fl in "" "--smooth-inputs"; do
    out=$(timeout 300 python bench.py --workload $wl --steps 300 --warmup 20 --no-cpu-baseline --no-e2e --no-train $fl 2>&1 | tail -1)
    echo "$wl inputs[$fl] $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"]*1e3,1), "us  frac", round(d["roofline"]["frac"],3))' 2>&1 | tail -1)" | tee -a gpurun_out/smooth_inputs.txt
  done
done
