#!/bin/bash
# build tuning variants of libsnde.so: scripts/build_variants.sh name "defs" [name "defs" ...]
set -e
cd "$(dirname "$0")/.."
mkdir -p self-supervised-normal-depth-estimation_b200/variants
while [ $# -gt 1 ]; do
  SNDE_LIB=$PWD/self-supervised-normal-depth-estimation_b200/variants/libsnde_$1.so SNDE_NVCC_DEFS="$2" python -m snde_b200.build --force -v 2>&1 | grep -A2 "ILb1ELb0ELb0ELb1ELb0ELb0" | grep -E "registers|spill" | tr '\n' ' '; echo " <- $1"
  shift 2
done
