#!/bin/bash
# Build an experimental variant of the C-ABI library for A/B timing (tools/prof_run.py with SEPAIHRD_LIB=...).
#   tools/build_variant.sh NAME [-DFLAG ...]   ->  tools/exp/libsepaihrd_NAME.so (+ .log with ptxas -v)
set -e
cd "$(dirname "$(readlink -f "$0")")/.."
name=$1; shift
CS=mathematical-modeling-of-infectious-diseases-v1_b200/csrc
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr \
  -I include "$@" -shared -o tools/exp/libsepaihrd_$name.so $CS/*.cu -lcudart > tools/exp/$name.log 2>&1
grep -A2 "sepaihrd_batch_kernelILi4ELb0ELi0E" tools/exp/$name.log | grep -E "spill|registers"
