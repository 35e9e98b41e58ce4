#!/usr/bin/env python
"""The reference's shipped phase-2 configuration at length: ONE Metropolis-Hastings chain, burn-in 5 000, adaptation every 100
iterations (data/configuration/mcmc_settings.txt; its 100 000 iterations shortened by --iterations), through the C++ host mirror on
the device objective, with look-ahead.  Reports iterations/s for the whole run and what share of the wall time the host spends
in the sampler itself (proposals, commits, rank-1 updates, the O(t P^2) covariance recomputation over the history).

    python tools/shipped_mcmc_run.py [--iterations 30000] [--sequential-iterations 6000]
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g

ap = argparse.ArgumentParser()
ap.add_argument("--iterations", type=int, default=30000)
ap.add_argument("--sequential-iterations", type=int, default=6000)
a = ap.parse_args()
pkg = g.load_package()
from sepaihrd_b200 import hostlib
p = pkg.load_default_problem()
x0 = p.base_params()
m = hostlib.HostModel(p)
m.calculate(x0)
st = dict(burn_in=5000, adaptation_period=100, n_chains=1, seed=3, store_samples=0, write_trace=0, write_checkpoints=0)
r = m.metropolis(dict(st, mcmc_iterations=a.iterations, lookahead=0), x0)
launch_ms = 0.57                      # one small launch through the host layer (profiles/r02_v17_small_batch_latency.txt)
out = dict(run="look-ahead", iterations=a.iterations, seconds=r["ms"] / 1e3, iterations_per_s=(a.iterations - 1) / r["ms"] * 1e3, launches=r["launches"],
           evaluations=r["evaluations"], acceptance_rate=r["acceptance_rate"], final_scale=r["final_scale"], best=r["best_value"],
           host_sampler_share_of_wall=max(0.0, 1.0 - r["launches"] * launch_ms / r["ms"]))
print(json.dumps(out), flush=True)
s = m.metropolis(dict(st, mcmc_iterations=a.sequential_iterations, lookahead=1), x0)
print(json.dumps(dict(run="sequential", iterations=a.sequential_iterations, seconds=s["ms"] / 1e3,
                      iterations_per_s=(a.sequential_iterations - 1) / s["ms"] * 1e3, launches=s["launches"])), flush=True)
