#!/usr/bin/env python
"""bench.py -- SEPAIHRD Dopri5 + Poisson likelihood evaluations per second on N B200s.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1      # CPU arm (oracle on host cores)

Workload (BASELINE.json configs[1]): a batched likelihood sweep of
B = 1,048,576 parameter sets PER GPU (weak scaling), 4 age groups, Spain-2020 window; set b =
clamp(base + sigma * N(0,1)) -- the reference benchmark's own jitter recipe
(sepaihrd_objective_benchmark_main.cpp:452-460), std::mt19937(1); all 1,048,576 sets are distinct draws.
One "step" = one pass of the hot path over the batch = ONE launch of the fused kernel.

value   : whole-job evals/s with the parameter batch resident in HBM; CUDA events on the launching
          stream, barrier + synchronize on both sides, max over ranks.
e2e     : the same metric through the host-buffer C-ABI call (sepaihrd_eval_batch): pinned host
          params H2D + kernel + logL/status D2H inside the timed region, every step.
roofline: FP64 pipe.  achieved = algorithmic FLOP per launch / average launch duration, with the
          algorithmic FLOP of SURVEY.md 8(d): sum_b attempts_b * 3750 + 22,000 * B (unfused count);
          peak = 2 * measured DFMA/s of this GPU (sepaihrd_measure_fp64_peak, run just before the
          timed region -- MEASURED_PEAKS.json has no FP64 entry).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

METRIC = "SEPAIHRD Dopri5+Poisson likelihood evals/sec"
UNIT = "evals/s"
B_PER_GPU = 1 << 20
DISTINCT = 1 << 20              # every set of the batch is its own draw
FLOP_PER_ATTEMPT = 3750.0      # 6 RHS + stage/solution/error combinations + error norm, n = 4
FLOP_PER_SET_FIXED = 22000.0   # 3672 likelihood terms * ~6
NCU_CAPTURE_FILE = os.path.join(ROOT, "profiles", "current_ncu_capture.json")
KERNEL_SOURCES = ("sepaihrd_kernels.cuh", "sepaihrd_constraints.cuh")


def kernel_source_hash() -> str:
    """sha256 over the kernel sources: ties a committed ncu capture to the build it was taken from."""
    import hashlib
    h = hashlib.sha256()
    for f in KERNEL_SOURCES:
        with open(os.path.join(entry.CSRC, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def ncu_capture(batch: int):
    """The numbers of the committed `ncu --set full` capture of the headline kernel (profiles/current_ncu_capture.json, written
    by tools/ncu_summary.py --json): DRAM traffic per launch, FP64-pipe utilisation, issue efficiency.  They are NOT measured by
    this run; `stale` says whether the kernel sources have changed since the capture."""
    try:
        with open(NCU_CAPTURE_FILE) as f:
            c = json.load(f)
    except Exception:
        return None
    c = dict(c)
    c["stale"] = c.get("kernel_source_hash") != kernel_source_hash()
    c["same_batch"] = int(c.get("sets_per_launch", -1)) == int(batch)
    return c


def workload_config(n_gpus: int, batch: int) -> dict:
    return {"workload": "batched likelihood sweep: 1M jittered parameter sets per GPU, 4 age groups, Spain-2020 "
                        "window (BASELINE configs[1])",
            "sets_per_gpu": batch, "global_sets": batch * n_gpus, "params_per_set": 62, "age_groups": 4,
            "output_days": 326, "tolerances": "abs=rel=1e-6", "param_distribution": "clamp(base+sigma*N(0,1)), mt19937(1)",
            "distinct_sets": min(DISTINCT, batch), "math": "fast (FMA)", "parallelism": f"independent shards x{n_gpus}",
            "ordering": "sets handed to the warps in the order of a fitted attempt-profile predictor (keys + counting sort inside every timed "
                        "step; the model fit, once per distribution, outside)",
            "l2": "inputs (520 MB of parameters per launch) larger than the 126 MB L2; two buffers alternate"}


def make_params(pkg, orc, batch: int):
    prob = pkg.load_default_problem()
    oracle = orc.Oracle(prob)
    distinct = oracle.jitter_params(min(DISTINCT, batch), seed=1)
    reps = (batch + len(distinct) - 1) // len(distinct)
    return prob, oracle, np.tile(distinct, (reps, 1))[:batch]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.flush()
        rows = []
        with open(self.f.name) as f:
            for line in f:
                parts = [x.strip() for x in line.split(",")]
                if len(parts) >= 7:
                    try:
                        rows.append((float(parts[0]), float(parts[1]), float(parts[2]), parts[3:7]))
                    except ValueError:
                        pass
        os.unlink(self.f.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = [r for r in rows if r[2] > 300.0] or rows     # samples under load (power well above idle)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in load for i in range(4) if r[3][i].lower().startswith("active")})
        return {"sm_mhz": float(np.median([r[0] for r in load])), "sm_max_mhz": rows[0][1],
                "power_w_max": max(r[2] for r in rows), "samples": len(rows), "samples_under_load": len(load),
                "reasons": reasons}


def host_cores() -> int:
    """Host threads the CPU arm may use: every core this process is allowed on.  Passed to the oracle EXPLICITLY because
    torchrun exports OMP_NUM_THREADS=1 to its ranks, which would silently shrink the CPU baseline to one core."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_baseline(oracle, params, target_seconds: float = 12.0):
    """The oracle (restated reference path, kind = "port") on all host cores, on a bounded sample."""
    nt = host_cores()
    probe = params[:256]
    t0 = time.perf_counter()
    _, _, _, used = oracle.eval_batch(probe, nthreads=nt)
    rate = len(probe) / (time.perf_counter() - t0)
    m = int(min(len(params), max(1024, rate * target_seconds)))
    t0 = time.perf_counter()
    ll, st, steps, used = oracle.eval_batch(params[:m], nthreads=nt)
    dt = time.perf_counter() - t0
    return {"value": m / dt, "unit": UNIT, "cores": int(used), "kind": "port",
            "sample": f"first {m} sets of the same batch, OpenMP schedule(dynamic) over sets, {dt:.1f} s; the oracle is the "
                      "dependency-free restatement of the reference path (the reference needs Boost/Eigen, absent here)"}, ll, st, steps


def run_reference(args, json_out):
    """--impl reference: the reference's CPU path (oracle port) on the host cores, same config/metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pkg = entry.load_package(); orc = entry.load_oracle()
    orc.build()
    prob, oracle, params = make_params(pkg, orc, B_PER_GPU)
    probe = params[:256]
    nt = host_cores()
    t0 = time.perf_counter(); oracle.eval_batch(probe, nthreads=nt); rate = 256 / (time.perf_counter() - t0)
    budget = 150.0 / max(1, args.steps + args.warmup)          # whole run within a few minutes
    m = int(min(len(params), max(512, rate * min(20.0, budget))))
    for _ in range(args.warmup):
        oracle.eval_batch(params[:m], nthreads=nt)
    t0 = time.perf_counter()
    used = 1
    for s in range(args.steps):
        off = (s * m) % (len(params) - m + 1)
        _, _, _, used = oracle.eval_batch(params[off:off + m], nthreads=nt)
    dt = time.perf_counter() - t0
    value = args.steps * m / dt
    sample = f"{m} sets per step (bounded sample of the 1M-set batch), {used} OpenMP threads"
    json_out.write(json.dumps({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                      "data": "synthetic", "config": workload_config(args.gpus, B_PER_GPU),
                      "cpu_baseline": {"value": value, "unit": UNIT, "cores": int(used), "kind": "port", "sample": sample},
                      "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}) + "\n")
    json_out.flush()


MH_LONG_ITERATIONS = 300
MH_CHAINS, MH_ITERATIONS = 4096, 30          # BASELINE configs[2]: 4096 seeded chains (strong scaling: the chains are split over the ranks)
PSO_PARTICLES, PSO_ITERATIONS = 65536, 30     # BASELINE configs[3]: 65,536 particles, global-best topology (strong scaling)


def _max_over_ranks(dist, world, dev, values):
    import torch
    t = torch.tensor(values, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.cpu()]


def _gather_bytes(dist, world, rank, dev, arr):
    """Rank 0 gets the list of every rank's array (uint8 / float64 numpy), the others None."""
    if world == 1:
        return [arr]
    out = [None] * world
    dist.all_gather_object(out, arr)
    return out if rank == 0 else None


def run_caller_configs(pkg, prob, rank, world, local_rank, dist):
    """BASELINE configs[2] (multi-chain Metropolis-Hastings) and configs[3] (particle swarm) with the callers RESIDENT ON THE
    DEVICES and their per-iteration collective on device buffers; strong-scaled over the ranks of this job; outside the headline
    timed region.  Each record carries wall time (max over ranks), the CUDA-event split of an iteration, the exchange transport
    and its parity gate: the seeded accept matrix / global-best trace of the sharded run equals the ONE-rank run bit for bit."""
    import torch
    from sepaihrd_b200 import drivers, resident
    from sepaihrd_b200.distributed import Comm
    from sepaihrd_b200.evaluator import BatchEvaluator
    dev = torch.device("cuda", local_rank)
    out = {}
    transports = ["p2p", "nccl"] if world > 1 else [None]
    base = prob.base_params()

    def summarise(r, n_iter, units):
        wall, setup = _max_over_ranks(dist, world, dev, [r["run_seconds"], r["setup_seconds"]])
        names = sorted(r["phase_seconds"])
        phases = _max_over_ranks(dist, world, dev, [r["phase_seconds"][k] for k in names])
        gpu = sum(phases)
        rec = {"wall_s": wall, "setup_s": setup, "iterations": n_iter, "ms_per_iteration": wall / n_iter * 1e3,
               "phase_ms_per_iteration": {k: v / n_iter * 1e3 for k, v in zip(names, phases)},
               "exchange_us_per_iteration": dict(zip(names, phases)).get("exchange", 0.0) / n_iter * 1e6,
               "host_share_of_wall": max(0.0, 1.0 - gpu / wall) if wall > 0 else None,
               "evals_per_s": units / wall, "transport": r["transport"], "fallback_reason": r["fallback_reason"],
               "exchange_status": r["exchange_status"]}
        return rec

    # ---- configs[2]: 4096 chains x 30 iterations, MCMC_REFLECT ----------------------------------------------------------
    try:
        rp = prob.__class__.from_json(dict(prob.to_json(), constraint_mode=1))
        with BatchEvaluator(rp, device=local_rank) as ev:
            ev.eval_batch(np.tile(base, (64, 1)))
            rec = {"workload": f"{MH_CHAINS} seeded Metropolis-Hastings chains x {MH_ITERATIONS} iterations, reflect mode, chains sharded over "
                               f"{world} rank(s); device-resident sampler (propose / accept kernels) + per-iteration all-gather of the log-likelihoods",
                   "scaling": "strong"}
            runs = {}
            for tr in transports:
                resident.run_mh_resident(ev, prob.sigmas, base, MH_CHAINS, 4, 1234, rank, world, transport=tr)       # warm-up (module load, IPC mapping)
                r = resident.run_mh_resident(ev, prob.sigmas, base, MH_CHAINS, MH_ITERATIONS, 1234, rank, world, transport=tr)
                runs[r["transport"]] = (r, summarise(r, MH_ITERATIONS - 1, MH_CHAINS * (MH_ITERATIONS - 1)))
            first = runs[next(iter(runs))][0]
            rec["transports"] = {k: v[1] for k, v in runs.items()}
            rec.update(runs[next(iter(runs))][1])
            parts = _gather_bytes(dist, world, rank, dev, first["accepts"])
            xs = _gather_bytes(dist, world, rank, dev, first["x"])
            if rank == 0:
                acc = np.concatenate(parts, axis=1)
                one = first if world == 1 else resident.run_mh_resident(ev, prob.sigmas, base, MH_CHAINS, MH_ITERATIONS, 1234, 0, 1)
                rec["parity"] = {"accept_rate": float(acc.mean()), "accept_checksum": int(np.packbits(acc).astype(np.int64).sum()),
                                 "accept_matrix_equals_one_rank_run": bool(np.array_equal(acc, one["accepts"])),
                                 "states_equal_one_rank_run": bool(np.array_equal(np.concatenate(xs), one["x"])),
                                 "best_trace_equals_one_rank_run": bool(np.array_equal(first["best_trace"], one["best_trace"])),
                                 "best_logpost": float(first["best_trace"][-1])}
            if world > 1:
                dist.barrier()
            # the same run in LOOK-AHEAD WINDOWS (sepaihrd_mh_window_*): every chain proposes its next K iterations at once, one
            # likelihood launch scores them, every chain commits up to its first accept; K from the shard size (4096 proposals per launch)
            try:
                tr0 = transports[0]
                resident.run_mh_resident(ev, prob.sigmas, base, MH_CHAINS, 4, 1234, rank, world, transport=tr0, lookahead=None)
                la = resident.run_mh_resident(ev, prob.sigmas, base, MH_CHAINS, MH_ITERATIONS, 1234, rank, world, transport=tr0, lookahead=None)
                la_rec = summarise(la, MH_ITERATIONS - 1, MH_CHAINS * (MH_ITERATIONS - 1))
                la_rec.update({"window_length": la["lookahead"], "windows": la["windows"], "ms_per_window": la_rec["wall_s"] / max(la["windows"], 1) * 1e3,
                               "proposals_scored_per_rank": la["evaluations"],
                               "speedup_over_one_iteration_per_launch": rec["wall_s"] / la_rec["wall_s"] if la_rec["wall_s"] > 0 else None})
                la_parts = _gather_bytes(dist, world, rank, dev, la["accepts"])
                la_xs = _gather_bytes(dist, world, rank, dev, la["x"])
                if rank == 0:
                    la_rec["parity"] = {"accept_matrix_equals_one_iteration_run": bool(np.array_equal(np.concatenate(la_parts, axis=1), acc)),
                                        "states_equal_one_iteration_run": bool(np.array_equal(np.concatenate(la_xs), np.concatenate(xs)))}
                # a window ends for the RUN when its slowest chain is done, so 30 iterations show little of it; 300 iterations beside them
                lk = resident.run_mh_resident(ev, prob.sigmas, base, MH_CHAINS, MH_LONG_ITERATIONS, 1234, rank, world, transport=tr0)
                lw = resident.run_mh_resident(ev, prob.sigmas, base, MH_CHAINS, MH_LONG_ITERATIONS, 1234, rank, world, transport=tr0, lookahead=None)
                w_lock, w_win = _max_over_ranks(dist, world, dev, [lk["run_seconds"], lw["run_seconds"]])
                same = float(np.array_equal(lk["accepts"], lw["accepts"]) and np.array_equal(lk["x"], lw["x"]))
                (all_same,) = _max_over_ranks(dist, world, dev, [1.0 - same])
                la_rec["long_run"] = {"iterations": MH_LONG_ITERATIONS, "one_iteration_per_launch_wall_s": w_lock, "lookahead_wall_s": w_win,
                                      "windows": lw["windows"], "speedup": w_lock / w_win if w_win > 0 else None,
                                      "chain_iterations_per_s": MH_CHAINS * (MH_LONG_ITERATIONS - 1) / w_win if w_win > 0 else None,
                                      "accept_matrix_and_states_equal_on_every_rank": all_same == 0.0}
                rec["lookahead"] = la_rec
            except Exception as exc:
                rec["lookahead"] = {"error": f"{type(exc).__name__}: {exc}"}
            if world > 1:
                dist.barrier()
            # the round-1 path beside it: host sampler (C++) + numpy -> H2D -> all_gather on a list of tensors -> .cpu() per rank
            comm = Comm()
            comm.barrier(); t0 = time.perf_counter()
            old = drivers.run_multichain_mh(ev.eval_batch, prob.sigmas, prob.lower_bound, prob.upper_bound, base, MH_CHAINS, MH_ITERATIONS,
                                            seed=1234, comm=comm, record_accepts=True)
            comm.barrier(); dt = time.perf_counter() - t0
            old_parts = _gather_bytes(dist, world, rank, dev, old["accepts"])
            w, e, c = _max_over_ranks(dist, world, dev, [dt, old["eval_seconds"], old["comm_seconds"]])
            rec["host_staged_path"] = {"wall_s": w, "ms_per_iteration": w / (MH_ITERATIONS - 1) * 1e3, "eval_ms_per_iteration": e / (MH_ITERATIONS - 1) * 1e3,
                                       "exchange_us_per_iteration": c / (MH_ITERATIONS - 1) * 1e6,
                                       "host_sampler_share_of_wall": max(0.0, 1.0 - (e + c) / w)}
            if rank == 0:
                rec["parity"]["accept_matrix_equals_host_sampler"] = bool(np.array_equal(np.concatenate(old_parts, axis=1), np.concatenate(parts, axis=1)))
            out["mh4096"] = rec
    except Exception as exc:                                   # never lose the headline line over a caller configuration
        out["mh4096"] = {"error": f"{type(exc).__name__}: {exc}"}

    # ---- configs[3]: 65,536 particles x 30 iterations, global best --------------------------------------------------------
    try:
        with BatchEvaluator(prob, device=local_rank) as ev:
            ev.eval_batch(np.tile(base, (64, 1)))
            rec = {"workload": f"particle swarm of {PSO_PARTICLES} particles x {PSO_ITERATIONS} iterations (STANDARD update, global-best topology), "
                               f"particles sharded over {world} rank(s); device-resident swarm + per-iteration all-gather of one 64-double record per rank",
                   "scaling": "strong"}
            runs = {}
            for tr in transports:
                resident.run_pso_resident(ev, PSO_PARTICLES, 2, 7, initial=base, rank=rank, world=world, transport=tr)
                r = resident.run_pso_resident(ev, PSO_PARTICLES, PSO_ITERATIONS, 7, initial=base, rank=rank, world=world, transport=tr)
                runs[r["transport"]] = (r, summarise(r, PSO_ITERATIONS + 1, PSO_PARTICLES * (PSO_ITERATIONS + 1)))
            first = runs[next(iter(runs))][0]
            rec["transports"] = {k: v[1] for k, v in runs.items()}
            rec.update(runs[next(iter(runs))][1])
            if rank == 0:
                one = first if world == 1 else resident.run_pso_resident(ev, PSO_PARTICLES, PSO_ITERATIONS, 7, initial=base, rank=0, world=1)
                rec["parity"] = {"global_best_trace_equals_one_rank_run": bool(np.array_equal(first["trace"], one["trace"])),
                                 "best_position_equals_one_rank_run": bool(np.array_equal(first["best_position"], one["best_position"])),
                                 "best_first": float(first["trace"][0]), "best_last": float(first["trace"][-1]),
                                 "transports_agree": bool(all(np.array_equal(v[0]["trace"], first["trace"]) for v in runs.values()))}
            if world > 1:
                dist.barrier()
            comm = Comm()
            comm.barrier(); t0 = time.perf_counter()
            old = drivers.run_pso(None, prob.sigmas, prob.lower_bound, prob.upper_bound, PSO_PARTICLES, PSO_ITERATIONS, seed=7, initial=base,
                                  comm=comm, device_ctx=ev.handle)
            comm.barrier(); dt = time.perf_counter() - t0
            w, e, c = _max_over_ranks(dist, world, dev, [dt, old["eval_seconds"], old["comm_seconds"]])
            rec["host_staged_path"] = {"wall_s": w, "ms_per_iteration": w / (PSO_ITERATIONS + 1) * 1e3, "eval_ms_per_iteration": e / (PSO_ITERATIONS + 1) * 1e3,
                                       "exchange_us_per_iteration": c / (PSO_ITERATIONS + 1) * 1e6}
            if rank == 0:
                rec["parity"]["trace_equals_host_staged_path"] = bool(np.array_equal(old["trace"], first["trace"]))
            out["pso65536"] = rec
    except Exception as exc:
        out["pso65536"] = {"error": f"{type(exc).__name__}: {exc}"}
    if world == 1:
        out["mh_one_chain"] = run_one_chain(pkg, prob)
        out["ppc100k"] = run_posterior_predictive(pkg, prob)
    return out


PPC_DRAWS = 100000


def _ppc_numpy(problem, oracle_obj, P, s0, probs):
    """ResultAggregator::aggregatePosteriorPredictives (.cpp:276-371) in numpy on ORACLE trajectories: daily increments clamped at 0,
    their running sums, exact linear-interpolation quantiles over the valid draws -> [6][T][n][Q]."""
    tr, st = oracle_obj.simulate_from_state(P, s0, what=1)          # [B][K][3n]: D | CumH | CumICU
    n = problem.n_ages
    T = int((problem.times >= 0).sum()); first = problem.n_times - T
    tr = tr[st == 0]
    blocks = (tr[:, :, n:2 * n], tr[:, :, 2 * n:3 * n], tr[:, :, 0:n])
    init = (s0[9 * n:10 * n], s0[10 * n:11 * n], s0[8 * n:9 * n])
    daily = []
    for X, x0 in zip(blocks, init):
        prev = X[:, first - 1] if first > 0 else np.broadcast_to(x0, X[:, 0].shape)
        daily.append(np.maximum(0.0, np.diff(np.concatenate([prev[:, None, :], X[:, first:]], axis=1), axis=1)))
    series = daily + [np.cumsum(d, axis=1) for d in daily]
    return np.stack([np.moveaxis(np.quantile(S, probs, axis=0), 0, -1) for S in series])


def run_posterior_predictive(pkg, prob):
    """BASELINE configs[4]: 100 000 posterior draws -> trajectories -> the six posterior-predictive series -> five quantiles per
    (series, day, age class), ONE C-ABI call from page-locked host draws (sepaihrd_posterior_predictive), for the Spain problem (4 age
    classes) and the synthetic 16-age-class variant.  Gate in the same job: the quantiles of a 384-draw subsample equal numpy's on
    oracle trajectories within 1e-8."""
    import torch
    from sepaihrd_b200.evaluator import BatchEvaluator
    orc = entry.load_oracle()
    out = {}
    probs = (0.025, 0.05, 0.5, 0.95, 0.975)
    for ages in (4, 16):
        try:
            pq = prob if ages == 4 else prob.expand_ages(ages // 4)
            oq = orc.Oracle(pq)
            base = oq.jitter_params(4096, seed=11)
            draws = torch.from_numpy(np.tile(base, ((PPC_DRAWS + 4095) // 4096, 1))[:PPC_DRAWS].copy()).pin_memory().numpy()
            s0 = pq.data_initial_state
            with BatchEvaluator(pq, device=torch.cuda.current_device()) as ev:
                ev.posterior_predictive(draws, s0, probs)                       # allocates the work buffers
                ts = []
                for _ in range(3):
                    t0 = time.perf_counter(); q, valid = ev.posterior_predictive(draws, s0, probs); ts.append(time.perf_counter() - t0)
                sub, _ = ev.posterior_predictive(draws[:384], s0, probs)
            ref = _ppc_numpy(pq, oq, draws[:384], s0, probs)
            err = float((np.abs(sub - ref) / np.maximum(np.abs(ref), 1e-6)).max())
            dt = sorted(ts)[1]
            out[f"{ages}_ages"] = {"draws": PPC_DRAWS, "valid": int(valid), "seconds": dt, "draws_per_s": PPC_DRAWS / dt,
                                   "quantile_cells": int(q.size), "h2d_bytes": int(draws.nbytes), "d2h_bytes": int(q.nbytes),
                                   "series_gbytes_on_device": 6 * q.shape[1] * q.shape[2] * PPC_DRAWS * 8 / 1e9,
                                   "parity": {"subsample_draws": 384, "max_rel_quantile_diff_vs_numpy_on_oracle_trajectories": err, "within_1e-8": err < 1e-8}}
        except Exception as exc:
            out[f"{ages}_ages"] = {"error": f"{type(exc).__name__}: {exc}"}
    return out


ONE_CHAIN_ITERATIONS, ONE_CHAIN_BURN_IN = 2000, 1000


def run_one_chain(pkg, prob):
    """The reference's SHIPPED phase 2: ONE Metropolis-Hastings chain of sequential iterations (data/configuration/mcmc_settings.txt;
    MetropolisHastingsSampler.cpp:283-384), through the C++ host mirror on the device objective -- one evaluation per launch (what a
    drop-in behind calculate() gives) against the look-ahead sampler (host/optimizers.cpp runLookahead: the next K iterations'
    proposals in ONE launch), with the gate that both runs end in the same state, log-posterior, scale and acceptance count, and one
    host core running the CPU oracle beside them.  2000 iterations, the second 1000 with covariance adaptation."""
    try:
        import __graft_entry__ as entry
        from sepaihrd_b200 import hostlib
        x0 = prob.base_params()
        m = hostlib.HostModel(prob)
        m.calculate(x0)
        st = dict(mcmc_iterations=ONE_CHAIN_ITERATIONS, burn_in=ONE_CHAIN_BURN_IN, adaptation_period=100, n_chains=1, seed=3, store_samples=0,
                  write_trace=0, write_checkpoints=0)
        seq = m.metropolis(dict(st, lookahead=1), x0)
        la = m.metropolis(dict(st, lookahead=0), x0)
        m.close()
        same = bool(np.array_equal(la["last"], seq["last"]) and la["last_logpost"] == seq["last_logpost"] and la["best_value"] == seq["best_value"]
                    and la["final_scale"] == seq["final_scale"] and la["acceptance_rate"] == seq["acceptance_rate"])
        orc = entry.load_oracle()
        o = orc.Oracle(prob)
        P = o.jitter_params(192, seed=9)
        o.eval_batch(P[:8], nthreads=1)
        t0 = time.perf_counter(); o.eval_batch(P, nthreads=1); dt = time.perf_counter() - t0
        n = ONE_CHAIN_ITERATIONS - 1
        return {"iterations": ONE_CHAIN_ITERATIONS, "burn_in": ONE_CHAIN_BURN_IN,
                "sequential": {"iterations_per_s": n / seq["ms"] * 1e3, "launches": seq["launches"], "evaluations": seq["evaluations"]},
                "lookahead": {"iterations_per_s": n / la["ms"] * 1e3, "launches": la["launches"], "evaluations": la["evaluations"],
                              "iterations_per_launch": n / max(la["launches"], 1)},
                "acceptance_rate": seq["acceptance_rate"],
                "one_host_core_oracle_evals_per_s": len(P) / dt,
                "parity": {"lookahead_chain_equals_sequential_chain": same}}
    except Exception as exc:
        return {"error": f"{type(exc).__name__}: {exc}"}


def _claim_stdout():
    """Keep file descriptor 1 for the ONE JSON line: everything else that writes to stdout while the benchmark runs
    (NCCL prints its version banner there, libraries may log) is sent to stderr.  Returns the stream for the JSON line."""
    sys.stdout.flush()
    out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return out


def main():
    json_out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="parameter sets per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ordering", action="store_true", help="hand the sets to the warps as given (no ordering pass in front of the launches)")
    ap.add_argument("--no-configs", action="store_true", help="skip BASELINE configs[2] / [3] (MH 4096 chains, PSO 65,536 particles)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args, json_out)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if rank == 0:
        entry.build_cuda()
    if world > 1:
        dist.barrier()
    pkg = entry.load_package(); orc = entry.load_oracle()
    from sepaihrd_b200.evaluator import BatchEvaluator, measure_fp64_peak

    B = args.batch
    prob, oracle, params = make_params(pkg, orc, B)
    # every rank evaluates its own shard: rank r rotates the batch so shards differ (independent units, no collective)
    params = np.roll(params, shift=rank * 7919, axis=0)
    dev = torch.device("cuda", local_rank)
    h_params = torch.from_numpy(params).pin_memory()
    d_params = [h_params.to(dev, non_blocking=True), torch.from_numpy(params[::-1].copy()).to(dev)]
    d_ll = torch.empty(B, dtype=torch.float64, device=dev)
    d_st = torch.empty(B, dtype=torch.int32, device=dev)
    d_steps = torch.empty((B, 2), dtype=torch.int32, device=dev)
    h_ll = torch.empty(B, dtype=torch.float64).pin_memory()
    h_st = torch.empty(B, dtype=torch.int32).pin_memory()

    ev = BatchEvaluator(prob, device=local_rank)
    stream = torch.cuda.current_stream(dev)
    ev.set_stream(stream.cuda_stream)
    P = prob.n_params

    def step(i: int, with_steps: bool = False):
        ev.eval_into(d_params[i & 1].data_ptr(), B, P, d_ll.data_ptr(), d_st.data_ptr(), d_steps.data_ptr() if with_steps else 0)

    # the ordering pass in front of large launches (csrc/sepaihrd_order.cu): its model is fitted once per distribution of
    # parameter sets (a 2048-set pilot + a small least-squares problem, ~15 ms), outside the timed regions; what it costs PER
    # LAUNCH (two linear predictors per set, a counting sort: ~0.3 ms per 1M sets) is inside every timed step below
    if not args.no_ordering:
        ev.fit_ordering(d_params[0])
    else:
        ev.set_ordering(False)
    peak_dfma = measure_fp64_peak(local_rank)
    # attempts per set (for the algorithmic FLOP count), measured once outside the timed region
    step(0, True); step(1, True); torch.cuda.synchronize()
    attempts_total = float(d_steps.sum().item())           # buffer 1 is buffer 0 reversed: same total
    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0, _ = ev.counters()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record(stream)
    for i in range(args.steps):
        step(i)
    e1.record(stream)
    sync_all()
    ms_kernel = e0.elapsed_time(e1)
    launches1, _ = ev.counters()

    # ---- end to end through the host-buffer C-ABI call ------------------------------------------------
    for _ in range(2):
        ev.eval_host_into(h_params.data_ptr(), B, P, h_ll.data_ptr(), h_st.data_ptr())
    f0 = torch.cuda.Event(enable_timing=True); f1 = torch.cuda.Event(enable_timing=True)
    sync_all()
    t_wall0 = time.perf_counter()
    f0.record(stream)
    for i in range(args.steps):
        ev.eval_host_into(h_params.data_ptr(), B, P, h_ll.data_ptr(), h_st.data_ptr())
    f1.record(stream)
    sync_all()
    ms_e2e = max(f0.elapsed_time(f1), (time.perf_counter() - t_wall0) * 1e3)
    launches2, _ = ev.counters()
    clocks = sampler.stop() if sampler else None
    checksum = float(h_ll.sum().item())

    if world > 1:
        t = torch.tensor([ms_kernel, ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_kernel, ms_e2e = t[0].item(), t[1].item()

    # H2D rate of every rank while ALL ranks copy at once (what the end-to-end path has to live with: at 4 and 8 GPUs the ranks
    # share the host's memory system, and a rate below 520 MB / kernel time per rank exposes copy time the chunking cannot hide)
    sync_all()
    g0 = torch.cuda.Event(enable_timing=True); g1 = torch.cuda.Event(enable_timing=True)
    g0.record(stream)
    d_params[1].copy_(h_params, non_blocking=True)
    g1.record(stream)
    sync_all()
    h2d_gbs = B * P * 8 / (g0.elapsed_time(g1) * 1e-3) / 1e9
    if world > 1:
        tt = torch.tensor([h2d_gbs, -h2d_gbs], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MIN)
        h2d_min, h2d_max = tt[0].item(), -tt[1].item()
    else:
        h2d_min = h2d_max = h2d_gbs

    caller_configs = None
    if not args.no_configs:
        # free the sweep's 1 GB of parameter buffers first; the callers below allocate their own state
        caller_configs = run_caller_configs(pkg, prob, rank, world, local_rank, dist if world > 1 else None)

    if rank == 0:
        value = world * args.steps * B / (ms_kernel * 1e-3)
        e2e_value = world * args.steps * B / (ms_e2e * 1e-3)
        flop_per_launch = attempts_total * FLOP_PER_ATTEMPT + FLOP_PER_SET_FIXED * B
        launch_s = ms_kernel * 1e-3 / args.steps
        achieved = flop_per_launch / launch_s / 1e12
        peak = 2.0 * peak_dfma / 1e12
        cap = ncu_capture(B)
        sm_ghz = ((clocks or {}).get("sm_mhz") or 1965.0) / 1e3
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms_kernel / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f64", "data": "synthetic", "config": workload_config(world, B),
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * P * 8, "d2h_bytes_per_step": B * 12,
                       "ms_per_step": ms_e2e / args.steps,
                       "h2d_gbs_per_rank_all_ranks_copying": {"min": h2d_min, "max": h2d_max,
                                                              "needed_to_hide_the_copy": B * P * 8 / (ms_kernel / args.steps * 1e-3) / 1e9}},
               "gpu_launches": int(launches1 - launches0),
               "gpu_launches_e2e": int(launches2 - launches1),
               "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                            "traffic": (cap["dram_bytes_per_launch"] if cap and cap["same_batch"] and not cap["stale"] else None),
                            "kernel": "sepaihrd_batch_kernel<4,fast,LL>",
                            "flop_per_launch": flop_per_launch, "attempts_per_set": attempts_total / B,
                            "launch_ms": launch_s * 1e3,
                            "peak_source": "measured live: sepaihrd_measure_fp64_peak (dependent DFMA chains), x2 FLOP per DFMA; "
                                           "MEASURED_PEAKS.json has no FP64 entry",
                            "hbm_bytes_per_launch_algorithmic": B * (P * 8 + 12),
                            # the datasheet-order denominator beside the measured one: 148 SMs x 64 FP64 lanes x 2 FLOP x SM clock
                            "frac_of_datasheet_peak": achieved / (148 * 64 * 2 * sm_ghz / 1e3),
                            "datasheet_peak": 148 * 64 * 2 * sm_ghz / 1e3, "datasheet_clock_ghz": sm_ghz,
                            # BASELINE.json's "% FP64 peak (ncu fp64-pipe utilisation and issue efficiency)": the committed ncu --set full
                            # capture of this kernel (NOT measured by this run; "stale" = kernel sources changed since the capture)
                            "ncu_capture": cap},
               "clocks": clocks, "logl_checksum": checksum}
        if caller_configs is not None:
            out["configs"] = caller_configs
        if world == 1:
            # SURVEY.md 8(d) names a second distribution: uniform in bounds (the PSO initialisation recipe), mt19937(2).  Not the
            # headline (warps idle more when their 8 sets need different attempt counts); reported beside it, outside the
            # timed region above.
            try:
                uni = torch.from_numpy(oracle.uniform_params(B, seed=2)).to(dev)
                if not args.no_ordering:
                    ev.fit_ordering(uni)                       # another distribution: refit (outside the timed region, like above)
                step_u = lambda: ev.eval_into(uni.data_ptr(), B, P, d_ll.data_ptr(), d_st.data_ptr(), d_steps.data_ptr())
                step_u(); torch.cuda.synchronize()
                att_u = float(d_steps.sum().item())
                u0 = torch.cuda.Event(enable_timing=True); u1 = torch.cuda.Event(enable_timing=True)
                u0.record(stream)
                for _ in range(3):
                    step_u()
                u1.record(stream); torch.cuda.synchronize()
                ms_u = u0.elapsed_time(u1) / 3
                out["second_distribution"] = {"param_distribution": "uniform in bounds, mt19937(2)", "value": B / (ms_u * 1e-3), "unit": UNIT,
                                              "ms_per_step": ms_u, "attempts_per_set": att_u / B, "ordering": ev.ordering_state()[0] and not args.no_ordering,
                                              "roofline_frac": (att_u * FLOP_PER_ATTEMPT + FLOP_PER_SET_FIXED * B) / (ms_u * 1e-3) / 1e12 / peak}
                del uni
            except Exception as exc:           # informational only: never lose the headline line over it
                out["second_distribution"] = {"error": str(exc)}
        if not args.no_cpu_baseline and world == 1:
            base, ll_cpu, st_cpu, steps_cpu = cpu_baseline(oracle, params)
            m = len(ll_cpu)
            rel = np.abs(h_ll.numpy()[:m] - ll_cpu) / np.abs(ll_cpu)
            base["max_rel_logl_diff_vs_gpu"] = float(rel.max())
            # parity gates in the same job: accept/reject path and status word of every set of the sample
            step(0, True); torch.cuda.synchronize()
            base["step_count_mismatches"] = int((d_steps[:m].cpu().numpy() != steps_cpu).any(axis=1).sum())
            base["status_mismatches"] = int((d_st[:m].cpu().numpy().astype(np.uint32) != st_cpu).sum())
            base["sets_compared"] = int(m)
            out["cpu_baseline"] = base
        json_out.write(json.dumps(out) + "\n")
        json_out.flush()
    ev.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
