#!/usr/bin/env python
"""Full-size parity run: N DISTINCT parameter sets per distribution (jitter and uniform-in-bounds), GPU (FAST and STRICT) against
the CPU oracle on all host cores.  Prints the statistics DESIGN.md section 2 cites.

    python tools/full_parity.py [--sets 1048576] [--strict-sets 131072]
"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g

ap = argparse.ArgumentParser()
ap.add_argument("--sets", type=int, default=1 << 20)
ap.add_argument("--strict-sets", type=int, default=1 << 17)
a = ap.parse_args()
pkg = g.load_package(); orc = g.load_oracle()
from sepaihrd_b200.evaluator import BatchEvaluator, MATH_FAST, MATH_STRICT
p = pkg.load_default_problem()
o = orc.Oracle(p)
print(f"host cores: {os.cpu_count()}")
for name, P in (("jitter  (mt19937(1))", o.jitter_params(a.sets, seed=1)), ("uniform (mt19937(2))", o.uniform_params(a.sets, seed=2))):
    assert len(np.unique(P[:, :8], axis=0)) == len(P)
    t0 = time.perf_counter()
    ll_ref, st_ref, steps_ref, _ = o.eval_batch(P)
    t_cpu = time.perf_counter() - t0
    for mname, mode, n in (("fast", MATH_FAST, a.sets), ("strict", MATH_STRICT, min(a.sets, a.strict_sets))):
        with BatchEvaluator(p, device=0, math=mode) as ev:
            t0 = time.perf_counter()
            ll, st, steps = ev.eval_batch(P[:n], return_steps=True)
            t_gpu = time.perf_counter() - t0
        ok = st_ref[:n] == 0
        rel = np.abs(ll[ok] - ll_ref[:n][ok]) / np.abs(ll_ref[:n][ok])
        print(f"{name} {mname:6s}: {n} distinct sets, status equal {bool((st == st_ref[:n]).all())} (failed sets {int((~ok).sum())}), "
              f"max rel logL err {rel.max():.3e}, 99.99th pct {np.quantile(rel, 0.9999):.3e}, bit-equal {int((ll == ll_ref[:n]).sum())}, "
              f"sets with a different (accepted, rejected) pair {int((steps != steps_ref[:n]).any(axis=1).sum())}, "
              f"attempts/set {steps_ref[:n].sum() / n:.1f}; oracle {len(P) / t_cpu:.0f} evals/s, GPU host call {n / t_gpu:.3e} evals/s")
