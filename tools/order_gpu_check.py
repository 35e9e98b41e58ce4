#!/usr/bin/env python
"""What would ordering a widely spread batch before the launch buy ON THE GPU?  ("Round 2".)

A pilot of 4096 uniform-in-bounds sets is evaluated on the CPU oracle with per-day attempt counts; the first principal component of
that profile is regressed on the (standardised) parameters; the 1M-set uniform batch of bench.py's `second_distribution` is then
sorted ON THE HOST by the predicted component and the kernel is timed on the batch as given and as sorted (same sets, same results).

    python tools/order_gpu_check.py
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as g

pkg = g.load_package(); orc = g.load_oracle()
from sepaihrd_b200.evaluator import BatchEvaluator
p = pkg.load_default_problem(); o = orc.Oracle(p)
M, B = 4096, 1 << 20
pilot = o.uniform_params(M, seed=99)
t0 = time.time()
A = np.array([o.eval_one(x, want_interval_steps=True)["interval_steps"].sum(axis=1) for x in pilot], dtype=np.float64)
X = A - A.mean(0); U, S, Vt = np.linalg.svd(X, full_matrices=False)
mu, sd = pilot.mean(0), pilot.std(0) + 1e-300
F = np.hstack([np.ones((M, 1)), (pilot - mu) / sd])
w, *_ = np.linalg.lstsq(F, U[:, :2] * S[:2], rcond=None)
print(f"pilot + fit: {time.time() - t0:.1f} s on the host")
for dist, P in (("uniform", o.uniform_params(B, seed=2)), ("jitter", o.jitter_params(B, seed=1))):
    pred = np.hstack([np.ones((B, 1)), (P - mu) / sd]) @ w
    order1 = np.argsort(pred[:, 0], kind="stable")
    r = np.argsort(np.argsort(pred[:, 0])) * 4096 // B
    order2 = np.lexsort((pred[:, 1], r))
    with BatchEvaluator(p, device=0) as ev:
        res = {}
        for name, order in (("as given", None), ("sorted by predicted PC1", order1), ("PC1 in 4096 bins, then PC2", order2)):
            x = torch.from_numpy(P if order is None else np.ascontiguousarray(P[order])).cuda()
            ll, st, steps = ev.eval_batch(x, return_steps=True); torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                ev.eval_batch(x)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            ll = ll.cpu().numpy()
            if order is not None:
                back = np.empty_like(ll); back[order] = ll; ll = back
            res[name] = ll
            print(f"{dist:8s} {name:32s}: {ms:8.3f} ms  ({B / ms * 1e3:.4e} evals/s)  attempts/set {float(steps.sum()) / B:.2f}  identical results: {bool(np.array_equal(ll, res['as given']))}", flush=True)
