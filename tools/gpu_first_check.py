#!/usr/bin/env python
"""First-light GPU check: parity vs oracle (strict + fast), FP64 peak, quick timing."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
pkg = g.load_package(); orc = g.load_oracle()
from sepaihrd_b200.evaluator import BatchEvaluator, measure_fp64_peak, MATH_STRICT, MATH_FAST
import torch

p = pkg.load_default_problem()
o = orc.Oracle(p)
out = {}
peak = measure_fp64_peak(0)
print("fp64 peak DFMA/s = %.4e  (%.2f TFLOP/s FMA-counted)" % (peak, 2 * peak / 1e12)); out["fp64_dfma_per_s"] = peak

for name, P in (("jitter", np.vstack([p.base_params()[None], o.jitter_params(2047, seed=1)])), ("uniform", o.uniform_params(2048, seed=2))):
    ll_ref, st_ref, steps_ref, _ = o.eval_batch(P)
    for mname, m in (("strict", MATH_STRICT), ("fast", MATH_FAST)):
        with BatchEvaluator(p, device=0, math=m) as ev:
            ll, st, steps = ev.eval_batch(P, return_steps=True)
        rel = np.abs(ll - ll_ref) / np.abs(ll_ref)
        mism = int((steps != steps_ref).any(axis=1).sum())
        print(f"{name:8s} {mname:6s}: max rel logL err {rel.max():.3e}  bit-equal {int((ll == ll_ref).sum())}/{len(ll)}  "
              f"step-count mismatches {mism}  status equal {bool((st == st_ref).all())}  ll[0]={ll[0]:.12e}")
        out[f"{name}_{mname}"] = dict(max_rel=float(rel.max()), step_mismatch=mism)

# trajectories
P = o.jitter_params(64, seed=5)
tr_ref, _ = o.simulate_batch(P)
for mname, m in (("strict", MATH_STRICT), ("fast", MATH_FAST)):
    with BatchEvaluator(p, device=0, math=m) as ev:
        tr, st = ev.simulate_batch(P)
    den = np.maximum(np.abs(tr_ref), 1e-300)
    rel = np.abs(tr - tr_ref) / np.maximum(np.abs(tr_ref), 1.0)
    print(f"traj {mname}: max |diff|/max(|ref|,1) = {rel.max():.3e}; bit-equal frac {np.mean(tr == tr_ref):.4f}")

# timing
for B in (1 << 16, 1 << 18, 1 << 20):
    P = o.jitter_params(4096, seed=9)
    P = np.tile(P, (B // 4096, 1))
    dP = torch.from_numpy(P).cuda()
    for mname, m in (("fast", MATH_FAST), ("strict", MATH_STRICT)):
        if m == MATH_STRICT and B > (1 << 18): continue
        with BatchEvaluator(p, device=0, math=m) as ev:
            ll, st, steps = ev.eval_batch(dP, return_steps=True); torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); ll, st = ev.eval_batch(dP); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            att = steps.sum().item()
            print(f"B={B} {mname}: {ms:.2f} ms  {B / ms * 1e3:.4e} evals/s  attempts/set {att / B:.1f}")
            out[f"time_{mname}_{B}"] = ms
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/first_check.json", "w"), indent=1)
