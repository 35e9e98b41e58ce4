#!/usr/bin/env python
"""Time the host-buffer entry point (pinned params in, logL/status out) for one chunking of the copy/compute overlap.

    SEPAIHRD_E2E_SPLIT=48,8 python tools/e2e_split.py [--B 1048576] [--steps 6]
"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as g

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=1 << 20)
ap.add_argument("--steps", type=int, default=6)
a = ap.parse_args()
pkg = g.load_package(); orc = g.load_oracle()
from sepaihrd_b200.evaluator import BatchEvaluator
p = pkg.load_default_problem()
P = orc.Oracle(p).jitter_params(a.B, seed=1)
h = torch.from_numpy(P).pin_memory()
ll = torch.empty(a.B, dtype=torch.float64).pin_memory()
st = torch.empty(a.B, dtype=torch.int32).pin_memory()
with BatchEvaluator(p, device=0) as ev:
    for _ in range(2):
        ev.eval_host_into(h.data_ptr(), a.B, P.shape[1], ll.data_ptr(), st.data_ptr())
    ts = []
    for _ in range(a.steps):
        t0 = time.perf_counter()
        ev.eval_host_into(h.data_ptr(), a.B, P.shape[1], ll.data_ptr(), st.data_ptr())
        ts.append((time.perf_counter() - t0) * 1e3)
print(f"split={os.environ.get('SEPAIHRD_E2E_SPLIT', 'default')}: min {min(ts):.3f} ms  median {sorted(ts)[len(ts) // 2]:.3f} ms  "
      f"({a.B / min(ts) * 1e3:.4e} evals/s)  checksum {float(ll.sum()):.12e}")
