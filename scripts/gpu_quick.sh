#!/bin/bash
# quick GPU check: GPU parity tests + short bench (no CPU baseline)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 500 --warmup 20 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit: $?"
cat gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err
