#!/usr/bin/env python
"""Where do device and oracle step counts part?  (VERDICT r1 "What's weak" 1)

Runs the problems of tests/test_gpu_parity.py that used to tolerate <= 5 % of sets with different (accepted, rejected)
counts -- off-grid breakpoints on a non-uniform grid, and the four fuzzed problems -- in STRICT and FAST, reports the
mismatching sets of each mode SEPARATELY and, with a diagnostic build of the library
(tools/build_variant.sh dbg -DSEPAIHRD_DEBUG_INTERVALS; SEPAIHRD_LIB=tools/exp/libsepaihrd_dbg.so), the first output
interval and the first attempt at which the two part, with (t, step, err) of both sides.

    SEPAIHRD_LIB=$PWD/tools/exp/libsepaihrd_dbg.so python tools/strict_parity_diag.py
"""
import copy, ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as g

pkg = g.load_package(); orc = g.load_oracle()
from sepaihrd_b200 import capi
from sepaihrd_b200.evaluator import BatchEvaluator, MATH_FAST, MATH_STRICT

problem = pkg.load_default_problem()
CAP = 2048


def off_grid_problem():
    p2 = copy.deepcopy(problem)
    p2.beta_end_times = np.array([13.4, 63.0, 84.25, 111.0, 183.7, 237.0, 305.0])
    p2.kappa_end_times = np.array([12.9, 63.0, 85.5, 111.0, 183.7, 240.1, 305.0])
    keep = np.r_[0:40, 40:326:2]
    p2.times = p2.times[keep] * 1.0
    off = int(np.argmax(p2.times >= 0))
    sel = keep[off:] - 20
    p2.obs_hosp, p2.obs_icu, p2.obs_deaths = p2.obs_hosp[sel], p2.obs_icu[sel], p2.obs_deaths[sel]
    o2 = orc.Oracle(p2)
    return "off-grid", p2, o2, o2.uniform_params(48, seed=6)


def fuzz_problem(seed):
    rng = np.random.default_rng(100 + seed)
    p2 = copy.deepcopy(problem)
    n_days = int(rng.integers(40, 120))
    start = float(rng.choice([-20.0, -7.5, 0.0]))
    steps_ = rng.choice([1.0, 1.0, 0.5, 2.0], size=n_days)
    times = start + np.concatenate([[0.0], np.cumsum(steps_)])
    p2.times = times
    T = int((times >= 0).sum())
    pick = rng.integers(0, problem.obs_hosp.shape[0] - 1, T)
    p2.obs_hosp, p2.obs_icu, p2.obs_deaths = problem.obs_hosp[pick].copy(), problem.obs_icu[pick].copy(), problem.obs_deaths[pick].copy()
    for arr in (p2.obs_hosp, p2.obs_icu, p2.obs_deaths):
        arr[rng.random(arr.shape) < 0.05] = np.nan
        arr[rng.random(arr.shape) < 0.03] = -1.0
    span = times[-1]
    nb, nk = len(problem.beta_end_times), len(problem.kappa_end_times)

    def breakpoints(k):
        pts = np.sort(rng.uniform(0.05 * span, 0.95 * span, k - 1))
        on_grid = rng.random(k - 1) < 0.5
        pts = np.where(on_grid, np.round(pts), pts)
        pts = np.maximum.accumulate(pts + 1e-3 * np.arange(k - 1))
        return np.concatenate([pts, [span + 10.0]])
    p2.beta_end_times, p2.kappa_end_times = breakpoints(nb), breakpoints(nk)
    p2.constraint_mode = int(rng.integers(0, 2))
    p2.abs_tol, p2.rel_tol = float(rng.choice([1e-6, 1e-7])), float(rng.choice([1e-6, 1e-5]))
    if seed % 2 == 1:
        p2 = p2.expand_ages(4)
    o2 = orc.Oracle(p2)
    P = np.vstack([o2.jitter_params(40, seed=seed + 50), o2.uniform_params(24, seed=seed + 60)])
    P += rng.standard_normal(P.shape) * p2.sigmas * (rng.random(P.shape) < 0.1) * 30
    return f"fuzz{seed}", p2, o2, P


def many_off_grid(seed, B=4096):
    """A larger sample on the off-grid problem so that a per-mode rate is measurable."""
    name, p2, o2, _ = off_grid_problem()
    return f"off-grid x{B}", p2, o2, np.vstack([o2.uniform_params(B // 2, seed=seed), o2.jitter_params(B // 2, seed=seed + 1)])


lib = capi.load_library()
debug = hasattr(lib, "sepaihrd_debug_set_trace")
print("library:", os.environ.get("SEPAIHRD_LIB", "(default)"), "| diagnostic build:", debug)

cases = [off_grid_problem()] + [fuzz_problem(s) for s in range(4)] + [many_off_grid(21)]
for name, p2, o2, P in cases:
    B = len(P); K = p2.n_times
    ll_ref, st_ref, steps_ref, _ = o2.eval_batch(P)
    print(f"== {name}: B={B} n={p2.n_ages} K={K} tol=({p2.abs_tol:g},{p2.rel_tol:g}) beta_end={np.round(p2.beta_end_times, 4).tolist()} kappa_end={np.round(p2.kappa_end_times, 4).tolist()}")
    for mode, mname in ((MATH_STRICT, "STRICT"), (MATH_FAST, "FAST")):
        with BatchEvaluator(p2, device=0, math=mode) as ev:
            dP = torch.from_numpy(np.ascontiguousarray(P)).cuda()
            d_ll = torch.empty(B, dtype=torch.float64, device="cuda")
            d_st = torch.empty(B, dtype=torch.int32, device="cuda")
            ev.set_stream(torch.cuda.current_stream().cuda_stream)
            if debug:
                d_steps = torch.zeros((B, K, 2), dtype=torch.int32, device="cuda")
                d_trace = torch.zeros((B, CAP, 3), dtype=torch.float64, device="cuda")
                lib.sepaihrd_debug_set_trace.argtypes = [C.c_void_p, C.c_void_p]
                lib.sepaihrd_debug_set_trace(ev.handle, C.c_void_p(d_trace.data_ptr()))
            else:
                d_steps = torch.zeros((B, 2), dtype=torch.int32, device="cuda")
            ev.eval_into(dP.data_ptr(), B, P.shape[1], d_ll.data_ptr(), d_st.data_ptr(), d_steps.data_ptr())
            torch.cuda.synchronize()
            ll = d_ll.cpu().numpy(); st = d_st.cpu().numpy().astype(np.uint32); steps = d_steps.cpu().numpy()
            if debug:
                lib.sepaihrd_debug_set_trace(ev.handle, C.c_void_p(0))
        ok = (st_ref == 0) & (st == 0)
        tot = steps[:, -1, :] if debug else steps
        mism = np.where(ok & (tot != steps_ref).any(axis=1))[0]
        rel = np.abs(ll[ok] - ll_ref[ok]) / np.abs(ll_ref[ok])
        print(f"  {mname}: status equal {bool((st == st_ref).all())}; max rel logL {rel.max():.3e}; step-count mismatches {len(mism)} of {int(ok.sum())}: sets {mism[:12].tolist()}")
        if not debug:
            continue
        for b in mism[:3]:
            r = o2.eval_one(P[b], want_interval_steps=True)
            cum_ref = np.vstack([[0, 0], np.cumsum(r["interval_steps"], axis=0)])          # totals at grid point idx
            dev = steps[b]
            bad = np.where((dev != cum_ref).any(axis=1))[0]
            i0 = int(bad[0]) - 1
            print(f"    set {b}: oracle {steps_ref[b].tolist()} device {tot[b].tolist()}; first differing interval [{p2.times[i0]:g}, {p2.times[i0 + 1]:g}]: "
                  f"oracle (acc,rej)={r['interval_steps'][i0].tolist()} device={(dev[i0 + 1] - dev[i0]).tolist()}")
            tr_ref, _ = o2.trace_one(P[b], CAP)
            tr_dev = d_trace[b].cpu().numpy()
            n0 = int(cum_ref[i0].sum())                      # attempts before the interval (equal on both sides)
            for a in range(max(0, n0 - 1), min(len(tr_ref), n0 + 12)):
                tag = " " if (tr_ref[a, 0] == tr_dev[a, 0] and tr_ref[a, 1] == tr_dev[a, 1]) else "*"
                print(f"      {tag} attempt {a}: oracle t={tr_ref[a, 0]!r} dt={tr_ref[a, 1]!r} err={tr_ref[a, 2]:.17g} | device t={tr_dev[a, 0]!r} dt={tr_dev[a, 1]!r} err={tr_dev[a, 2]:.17g}")
