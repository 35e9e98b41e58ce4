#!/bin/bash
# A/B timing of experimental builds (tools/build_variant.sh) in one GPU session: tools/ab.sh name1 name2 ...
# Prints per variant the jitter and uniform 1M-set launch times and the logL checksum (must agree between variants).
cd "$(dirname "$(readlink -f "$0")")/.."
mkdir -p gpurun_out
for rep in 1 2; do
for v in "$@"; do
  for dist in jitter uniform; do
    echo -n "$v rep$rep: "
    SEPAIHRD_LIB=$PWD/tools/exp/libsepaihrd_$v.so python tools/prof_run.py --B 1048576 --launches 3 --dist $dist 2>&1 | tail -1
  done
done
done | tee gpurun_out/ab_$(date +%H%M%S).log
