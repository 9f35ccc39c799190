#!/bin/bash
# quick GPU check: micro-benchmarks + GPU parity tests
set -u
mkdir -p gpurun_out
./scripts/ubench/f32x2 > gpurun_out/ubench_f32x2.txt 2>&1; cat gpurun_out/ubench_f32x2.txt
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
