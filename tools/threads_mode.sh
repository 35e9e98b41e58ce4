#!/bin/bash
# throughput of calculate() called from an OpenMP loop (the unchanged reference pattern) for several thread counts
set -e
T=$(mktemp -d)
python - <<PY
import sys; sys.path.insert(0, ".")
import __graft_entry__ as e
pkg = e.load_package()
from sepaihrd_b200 import config
config.write_reference_tree(pkg.load_default_problem(), "$T")
PY
EXE=mathematical-modeling-of-infectious-diseases-v1_b200/host/sepaihrd_objective_benchmark
for th in 1 16 128 1024 4096; do
  n=$(( th * 16 )); [ $n -lt 256 ] && n=256
  $EXE --project-root $T --mode threads --threads $th --jitters $n --seed 1 | grep -A3 "OpenMP loop"
done
