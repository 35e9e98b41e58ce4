#!/usr/bin/env python
"""Latency of the host-buffer call (pinned params in, logL / status out, synchronous) and of the device-pointer launch for
small batches: what single-chain samplers, line searches and a few thousand chains per GPU see.

    python tools/small_batch_latency.py
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
pkg = g.load_package(); orc = g.load_oracle()
from sepaihrd_b200.evaluator import BatchEvaluator
p = pkg.load_default_problem()
P = orc.Oracle(p).jitter_params(16384, seed=1)
with BatchEvaluator(p, device=0) as ev:
    for B in (1, 8, 64, 512, 1024, 4096, 8192, 16384):
        h = torch.from_numpy(P[:B].copy()).pin_memory(); ll = torch.empty(B, dtype=torch.float64).pin_memory(); st = torch.empty(B, dtype=torch.int32).pin_memory()
        for _ in range(20): ev.eval_host_into(h.data_ptr(), B, 62, ll.data_ptr(), st.data_ptr())
        ts = []
        for _ in range(200):
            t0 = time.perf_counter(); ev.eval_host_into(h.data_ptr(), B, 62, ll.data_ptr(), st.data_ptr()); ts.append(time.perf_counter() - t0)
        ts.sort()
        d = h.cuda(); dl = torch.empty(B, dtype=torch.float64, device="cuda"); ds = torch.empty(B, dtype=torch.int32, device="cuda")
        ev.set_stream(torch.cuda.current_stream().cuda_stream)
        for _ in range(5): ev.eval_into(d.data_ptr(), B, 62, dl.data_ptr(), ds.data_ptr())
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): ev.eval_into(d.data_ptr(), B, 62, dl.data_ptr(), ds.data_ptr())
        e1.record(); torch.cuda.synchronize()
        print(f"B={B:6d}: host call median {ts[100]*1e3:.4f} ms  min {ts[0]*1e3:.4f} ms  ({B / ts[100]:.0f} evals/s) | device launch {e0.elapsed_time(e1) / 50:.4f} ms")
