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
ap.add_argument("--ages", type=int, default=4)
ap.add_argument("--traj", action="store_true", help="time the trajectory kernel (full state, device output)")
ap.add_argument("--host", action="store_true", help="time the host-buffer entry point (numpy in / numpy out)")
a = ap.parse_args()
pkg = g.load_package(); orc = g.load_oracle()
from sepaihrd_b200.evaluator import BatchEvaluator, MATH_FAST, MATH_STRICT
p = pkg.load_default_problem()
if a.ages != 4:
    p = p.expand_ages(a.ages // 4)
o = orc.Oracle(p)
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
    if a.traj:
        from sepaihrd_b200 import capi
        W = p.state_size; K = p.n_times
        d_out = torch.empty((a.B, K, W), dtype=torch.float64, device="cuda")
        d_st = torch.empty(a.B, dtype=torch.int32, device="cuda")
        ev.set_stream(torch.cuda.current_stream().cuda_stream)
        def tl():
            capi.check(ev._lib.sepaihrd_simulate_batch_device(ev._h, dP.data_ptr(), a.B, dP.shape[1], 0, 1, d_out.data_ptr(), d_st.data_ptr()))
        tl(); torch.cuda.synchronize()
        f0 = torch.cuda.Event(enable_timing=True); f1 = torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(a.launches):
            tl()
        f1.record(); torch.cuda.synchronize()
        tms = f0.elapsed_time(f1) / a.launches
        print(f"traj kernel: {tms:.3f} ms/launch  {a.B / tms * 1e3:.4e} draws/s  {a.B * K * W * 8 / tms / 1e6:.1f} GB/s written")
    if a.host:
        import time
        ev.eval_batch(P)
        t0 = time.perf_counter()
        for _ in range(a.launches):
            llh, sth = ev.eval_batch(P)
        print(f"host path: {(time.perf_counter() - t0) / a.launches * 1e3:.3f} ms/call")
    print(f"B={a.B} ages={a.ages} {a.dist} {a.math}: {ms:.3f} ms/launch  {a.B / ms * 1e3:.4e} evals/s  checksum {float(ll.sum()):.12e}")
