#!/usr/bin/env python
"""The reference's shipped phase-2 run -- ONE Metropolis-Hastings chain, sequential iterations
(data/configuration/mcmc_settings.txt; MetropolisHastingsSampler.cpp:283-384) -- through the C++ host mirror on the device
objective: sequential (one evaluation per launch, what a drop-in of calculate() gives) against the look-ahead sampler
(setting `lookahead`: the next K iterations' proposals in one launch, bit-identical chain), with one host core running the
CPU oracle beside them (the reference's own arithmetic for one calculate()).

    python tools/single_chain_mh.py [--iterations 3000] [--burn-in 1000] [--json out.json]
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g

ap = argparse.ArgumentParser()
ap.add_argument("--iterations", type=int, default=3000)
ap.add_argument("--burn-in", type=int, default=1000)
ap.add_argument("--seed", type=int, default=3)
ap.add_argument("--json", default="")
a = ap.parse_args()

pkg = g.load_package(); orc = g.load_oracle()
from sepaihrd_b200 import hostlib
p = pkg.load_default_problem()
x0 = p.base_params()
m = hostlib.HostModel(p)
m.calculate(x0)                                                   # context, first launch
st = dict(mcmc_iterations=a.iterations, burn_in=a.burn_in, adaptation_period=100, n_chains=1, seed=a.seed, store_samples=0,
          write_trace=0, write_checkpoints=0)
rows = []
ref = None
for name, la in (("sequential", 1), ("lookahead auto", 0), ("lookahead 8", 8), ("lookahead 32", 32), ("lookahead 128", 128)):
    r = m.metropolis(dict(st, lookahead=la), x0)
    if ref is None:
        ref = r
    same = bool(np.array_equal(r["last"], ref["last"]) and r["last_logpost"] == ref["last_logpost"] and r["best_value"] == ref["best_value"]
                and r["final_scale"] == ref["final_scale"] and r["acceptance_rate"] == ref["acceptance_rate"])
    rows.append(dict(run=name, lookahead=la, ms=r["ms"], iterations_per_s=(a.iterations - 1) / r["ms"] * 1e3, launches=r["launches"],
                     evaluations=r["evaluations"], iterations_per_launch=(a.iterations - 1) / max(r["launches"] - 0, 1),
                     acceptance_rate=r["acceptance_rate"], final_scale=r["final_scale"], best=r["best_value"], last_logpost=r["last_logpost"],
                     identical_to_sequential=same))
    print(json.dumps(rows[-1]), flush=True)
# a few chains: every chain looks ahead on its own (a launch holds at most 4096 proposals)
multi = []
for nc in (4, 64, 512):
    stc = dict(st, n_chains=nc, mcmc_iterations=600, burn_in=300)
    a0 = m.metropolis(dict(stc, lookahead=1), x0)
    a1 = m.metropolis(dict(stc, lookahead=0), x0)
    same = bool(np.array_equal(a0["last"], a1["last"]) and a0["best_value"] == a1["best_value"] and a0["acceptance_rate"] == a1["acceptance_rate"])
    multi.append(dict(chains=nc, iterations=600, lockstep_chain_iterations_per_s=nc * 599 / a0["ms"] * 1e3, lookahead_chain_iterations_per_s=nc * 599 / a1["ms"] * 1e3,
                      lockstep_launches=a0["launches"], lookahead_launches=a1["launches"], lookahead_evaluations=a1["evaluations"], identical=same))
    print(json.dumps(multi[-1]), flush=True)
# one host core, the oracle (== the reference's arithmetic for one calculate(); BASELINE.md: the port is faster than the reference build)
o = orc.Oracle(p)
P = o.jitter_params(256, seed=9)
o.eval_batch(P[:16], nthreads=1)
t0 = time.perf_counter(); o.eval_batch(P, nthreads=1); dt = time.perf_counter() - t0
cpu = dict(run="one host core, CPU oracle", evals_per_s=len(P) / dt, sample="256 jittered sets, 1 thread")
print(json.dumps(cpu), flush=True)
assert all(r["identical_to_sequential"] for r in rows), "look-ahead chain differs from the sequential chain"
assert all(r["identical"] for r in multi), "multi-chain look-ahead differs from the lockstep run"
if a.json:
    with open(a.json, "w") as f:
        json.dump(dict(runs=rows, multi=multi, cpu=cpu, iterations=a.iterations, burn_in=a.burn_in), f, indent=1)
