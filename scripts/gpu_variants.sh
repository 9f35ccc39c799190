#!/bin/bash
# bench every tuning variant under variants/ (short runs, device-resident number only)
mkdir -p gpurun_out; : > gpurun_out/variants.txt
for lib in self-supervised-normal-depth-estimation_b200/variants/libsnde_*.so; do
  out=$(SNDE_LIB=$PWD/$lib timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --no-e2e --no-train 2>&1 | tail -1)
  echo "$(basename $lib) $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"]*1e3,1), "us  loss", d["config"]["loss_check"])' 2>&1 | tail -1)" | tee -a gpurun_out/variants.txt
done
