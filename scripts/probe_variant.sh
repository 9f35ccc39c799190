#!/bin/bash
# compile snde_fused.cu alone with extra defines and print registers / spills / loop statistics of the hot instance
# usage: scripts/probe_variant.sh name "-DFOO=1 ..."
cd "$(dirname "$0")/.."
P=self-supervised-normal-depth-estimation_b200
mkdir -p /tmp/probe
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-ffp-contract=off,-fvisibility=hidden -Xptxas -v \
  -I include -I $P/csrc -DSNDE_BUILDING=1 $2 -fmad=true -c -o /tmp/probe/$1.o $P/csrc/snde_fused.cu 2> /tmp/probe/$1.log
grep -A2 "ILb1ELb0ELb0ELb1ELb0ELb0" /tmp/probe/$1.log | grep -E "registers|spill"
python scripts/sass_roles.py /tmp/probe/$1.o _Z17loss_strip_kernelILb1ELb0ELb0ELb1ELb0ELb0EEvN4snde7DevLossE | awk '/^instructions/ || /^row loop/ {print} /^  / {print substr($0,1,400)}' | awk '!seen[$0]++'
