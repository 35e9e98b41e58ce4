#!/usr/bin/env python
"""The warp-pair kernel (MATH_FAST_SPLIT, csrc/sepaihrd_split.cuh) against the FAST kernel: bit-identical results on jittered and
uniform-in-bounds sets, then launch times of both across batch sizes (CUDA events, device-resident inputs).

    tools/build_variant.sh split -DSEPAIHRD_WITH_SPLIT
    SEPAIHRD_LIB=$PWD/tools/exp/libsepaihrd_split.so python tools/split_check.py [--big]
(the shipped library does not contain the experiment: profiles/r02_split_kernel_experiment.txt has the measurement)
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as g

pkg = g.load_package(); orc = g.load_oracle()
from sepaihrd_b200.evaluator import BatchEvaluator, MATH_FAST, MATH_FAST_SPLIT
p = pkg.load_default_problem()
o = orc.Oracle(p)
big = "--big" in sys.argv
N = 1 << 20 if big else 1 << 16
P = np.vstack([o.jitter_params(N // 2, seed=1), o.uniform_params(N // 2, seed=2)])
P[7] = p.upper_bound * 50.0           # far outside: clamped
dP = torch.from_numpy(P).cuda()
with BatchEvaluator(p, device=0, math=MATH_FAST) as a, BatchEvaluator(p, device=0, math=MATH_FAST_SPLIT) as b:
    for B in (1, 5, 8, 9, 333, len(P)):
        la, sa, ta = a.eval_batch(dP[:B], return_steps=True)
        lb, sb, tb = b.eval_batch(dP[:B], return_steps=True)
        torch.cuda.synchronize()
        same = bool(torch.equal(la, lb) and torch.equal(sa, sb) and torch.equal(ta, tb))
        print(f"B={B}: bit-identical logL / status / steps: {same}", flush=True)
        if not same:
            d = (la != lb).nonzero().flatten()
            print("   differing sets:", d[:10].tolist(), "of", int(d.numel()), "| step diffs:", int((ta != tb).any(dim=1).sum()),
                  "| max rel", float(((la - lb).abs() / la.abs()).max()))
    def timeit(ev, x, reps):
        ev.eval_batch(x); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            ev.eval_batch(x)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    for dist, lo in (("jitter", 0), ("uniform", N // 2)):
        for B in (1, 64, 512, 1024, 4096, 8192, 16384, 65536 // 2, N // 2):
            x = dP[lo:lo + B]
            reps = 20 if B <= 65536 else 3
            tf, ts = timeit(a, x, reps), timeit(b, x, reps)
            print(f"{dist:8s} B={B:8d}: FAST {tf:9.4f} ms   SPLIT {ts:9.4f} ms   ratio {tf / ts:5.2f}x   ({B / ts * 1e3:.4e} evals/s split)", flush=True)
