#!/usr/bin/env python
"""Tiny driver for ncu captures: W warm-up launches + N timed launches of the fused kernel on B jittered sets.

    python tools/prof_run.py [--B 262144] [--dist jitter|uniform] [--math fast|strict] [--launches 1]
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as g

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=262144)
ap.add_argument("--dist", default="jitter")
ap.add_argument("--math", default="fast")
ap.add_argument("--launches", type=int, default=1)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--loop", type=int, default=6)
a = ap.parse_args()
pkg = g.load_package(); orc = g.load_oracle()
from sepaihrd_b200.evaluator import BatchEvaluator, MATH_FAST, MATH_STRICT
p = pkg.load_default_problem(); o = orc.Oracle(p)
D = min(a.B, 65536)
P = o.jitter_params(D, seed=1) if a.dist == "jitter" else o.uniform_params(D, seed=2)
P = np.tile(P, ((a.B + D - 1) // D, 1))[:a.B]
dP = torch.from_numpy(P).cuda()
with BatchEvaluator(p, device=0, math=MATH_FAST if a.math == "fast" else MATH_STRICT) as ev:
    for _ in range(a.warmup):
        ev.eval_batch(dP)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.launches):
        ll, st = ev.eval_batch(dP)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.launches
    print(f"B={a.B} {a.dist} {a.math} loop{a.loop}: {ms:.3f} ms/launch  {a.B / ms * 1e3:.4e} evals/s  checksum {float(ll.sum()):.12e}")
