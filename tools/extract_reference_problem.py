#!/usr/bin/env python
"""Extract the Spain-2020 problem constants from a reference checkout into the committed fixture.

    python tools/extract_reference_problem.py [/root/reference]

Reads ONLY data files of the reference (data/configuration/*.txt, data/contacts.csv,
data/processed/processed_data.csv) with this repo's own readers (sepaihrd_b200.config) and writes
mathematical-modeling-of-infectious-diseases-v1_b200/data/spain2020_problem.json.
/root/reference does not exist on the GPU box, so tests and bench.py use the JSON.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
prob = pkg.config.problem_from_reference_tree(root)
out = pkg.default_problem_path()
prob.save(out)
print(f"wrote {out}: n_ages={prob.n_ages} K={prob.n_times} n_obs={prob.n_obs} P={prob.n_params} "
      f"slots={prob.layout.count} bytes={os.path.getsize(out)}")
