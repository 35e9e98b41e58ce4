"""Sharded drivers of the batched calibrators (BASELINE.json configs[2] and [3]).

The sampler logic is the C++ host layer (host/optimizers.cpp: the reference's MetropolisHastingsSampler and
ParticleSwarmOptimization, batched); a driver here owns one shard of chains / particles per rank, evaluates its
shard with one fused-kernel launch per iteration and exchanges the few numbers the algorithm needs across ranks
(distributed.Comm: NCCL on the GPU box, gloo in the CPU tests).

``evaluate`` is any callable [B, P] -> [B] log-likelihoods: ``BatchEvaluator.eval_batch`` in production, the CPU
oracle in the parity tests.  Per-chain / per-particle random streams are indexed by GLOBAL chain / particle
number, so the visited states do not depend on the number of ranks.
"""
from __future__ import annotations

import time
from typing import Callable, Dict, Optional

import numpy as np

from . import hostlib
from .distributed import Comm, shard_range

Evaluate = Callable[[np.ndarray], np.ndarray]


def _host_threads(comm: Comm) -> None:
    """Give the C++ sampler loops this rank's share of the host cores (torchrun exports OMP_NUM_THREADS=1)."""
    import os
    if os.environ.get("SEPAIHRD_HOST_THREADS"):
        hostlib.set_threads(int(os.environ["SEPAIHRD_HOST_THREADS"]))
    elif comm.world > 1:
        hostlib.set_threads(max(1, (os.cpu_count() or 1) // comm.world))


def _loglik(evaluate: Evaluate, x: np.ndarray) -> np.ndarray:
    out = evaluate(x)
    if isinstance(out, tuple):          # BatchEvaluator.eval_batch returns (ll, status)
        out = out[0]
    return np.asarray(out, dtype=np.float64)


def run_multichain_mh(evaluate: Evaluate, sigmas, lower, upper, initial, n_chains: int, iterations: int, seed: int,
                      comm: Optional[Comm] = None, settings: Optional[Dict[str, float]] = None, record_accepts: bool = True):
    """``n_chains`` independent adaptive-Metropolis chains (MetropolisHastingsSampler.cpp:201-412 each), sharded over
    the ranks of ``comm``; one batch evaluation per iteration per rank; all_gather of the log-likelihoods per
    iteration.  Returns a dict with the gathered final state, the accept matrix of the local shard and timings."""
    comm = comm or Comm()
    _host_threads(comm)
    lo, hi = shard_range(n_chains, comm.rank, comm.world)
    counts = [shard_range(n_chains, r, comm.world)[1] - shard_range(n_chains, r, comm.world)[0] for r in range(comm.world)]
    pm = hostlib.ParameterManager(sigmas, lower, upper, mode=1)        # MCMC_REFLECT (MetropolisHastingsSampler.cpp:207-210)
    st = dict(mcmc_iterations=iterations, burn_in=iterations, n_chains=hi - lo, chain_offset=lo, seed=seed, store_samples=0)
    st.update(settings or {})
    st["n_chains"], st["chain_offset"] = hi - lo, lo
    mh = hostlib.MultiChainMH(pm, st)
    x0 = np.ascontiguousarray(initial, dtype=np.float64)
    lp0 = _loglik(evaluate, x0[None, :])[0]
    mh.begin(x0, np.full(hi - lo, lp0))
    accepts = np.zeros((iterations - 1, hi - lo), dtype=np.uint8) if record_accepts else None
    t_eval = t_comm = 0.0
    gathered = None
    best_trace = []
    while not mh.done:
        it = mh.iteration
        prop = mh.propose()
        t0 = time.perf_counter()
        lp = _loglik(evaluate, prop)
        t1 = time.perf_counter()
        acc = mh.accept(lp)
        if accepts is not None:
            accepts[it - 1] = acc
        cur_lp = mh.state()[1]
        t2 = time.perf_counter()
        gathered = comm.all_gather_varlen(cur_lp, counts)               # every rank sees every chain's likelihood
        t3 = time.perf_counter()
        best_trace.append(float(gathered.max()))
        t_eval += t1 - t0
        t_comm += t3 - t2
    x, lp, scale, n_acc = mh.state()
    return dict(rank=comm.rank, world=comm.world, chains=(lo, hi), x=x, logpost=lp, scale=scale, accepted=n_acc, accepts=accepts,
                all_logpost=gathered, best_trace=np.array(best_trace), eval_seconds=t_eval, comm_seconds=t_comm,
                evaluations=(iterations - 1) * (hi - lo) + 1)


def run_pso(evaluate: Evaluate, sigmas, lower, upper, swarm_size: int, iterations: int, seed: int, initial=None,
            comm: Optional[Comm] = None, settings: Optional[Dict[str, float]] = None, device_ctx=None, return_positions: bool = False):
    """Particle swarm with the global-best topology (ParticleSwarmOptimizer.cpp:106-247, 330-425, 576-618), particles
    sharded over the ranks; per iteration one batch evaluation per rank, then the global-best reduction.

    device_ctx (BatchEvaluator.handle): keep this rank's shard in the HBM of that evaluator's GPU (sepaihrd_swarm_*):
    per iteration one seed per particle goes down and one (value, index, position) triple comes back; `evaluate` is not
    called.  The visited positions are identical to the host-resident run."""
    comm = comm or Comm()
    if device_ctx is not None:
        return _run_pso_device(device_ctx, sigmas, lower, upper, swarm_size, iterations, seed, initial, comm, settings, return_positions)
    _host_threads(comm)
    lo, hi = shard_range(swarm_size, comm.rank, comm.world)
    pm = hostlib.ParameterManager(sigmas, lower, upper, mode=0)        # OPTIMIZATION_CLAMP (ModelCalibrator.cpp:62-66)
    st = dict(iterations=iterations, swarm_size=swarm_size, particle_offset=lo, local_count=hi - lo, seed=seed)
    st.update(settings or {})
    st["particle_offset"], st["local_count"] = lo, hi - lo
    sw = hostlib.Swarm(pm, st)
    sw.begin(initial)
    t_eval = t_comm = 0.0
    trace = []

    def evaluate_and_reduce():
        nonlocal t_eval, t_comm
        t0 = time.perf_counter()
        fit = _loglik(evaluate, sw.positions())
        t1 = time.perf_counter()
        v, i, pos = sw.tell(fit)
        gv, gi, gpos = comm.argmax_and_fetch(v, lo + i if i >= 0 else -1, pos)
        sw.set_global_best(gv, gpos)
        t2 = time.perf_counter()
        t_eval += t1 - t0
        t_comm += t2 - t1
        trace.append(sw.global_best()[0])

    evaluate_and_reduce()
    for it in range(iterations):
        sw.step(it)
        evaluate_and_reduce()
    val, pos = sw.global_best()
    return dict(rank=comm.rank, world=comm.world, particles=(lo, hi), best_value=val, best_position=pos, trace=np.array(trace),
                eval_seconds=t_eval, comm_seconds=t_comm, evaluations=(iterations + 1) * (hi - lo))


def _run_pso_device(device_ctx, sigmas, lower, upper, swarm_size, iterations, seed, initial, comm, settings, return_positions):
    _host_threads(comm)
    lo, hi = shard_range(swarm_size, comm.rank, comm.world)
    pm = hostlib.ParameterManager(sigmas, lower, upper, mode=0)
    st = dict(iterations=iterations, swarm_size=swarm_size, seed=seed)
    st.update(settings or {})
    st["particle_offset"], st["local_count"] = lo, hi - lo
    t_start = time.perf_counter()
    sw = hostlib.Swarm(pm, st)
    sw.begin_device(device_ctx, initial)          # allocates the swarm in HBM (+ pinned staging) and draws it
    t_setup = time.perf_counter() - t_start
    t_eval = t_comm = t_step = 0.0
    trace = []

    def evaluate_and_reduce():
        nonlocal t_eval, t_comm
        t0 = time.perf_counter()
        v, i, pos = sw.evaluate_device()
        t1 = time.perf_counter()
        gv, gi, gpos = comm.argmax_and_fetch(v, lo + i if i >= 0 else -1, pos)
        sw.set_global_best(gv, gpos)
        t2 = time.perf_counter()
        t_eval += t1 - t0
        t_comm += t2 - t1
        trace.append(sw.global_best()[0])

    evaluate_and_reduce()
    for it in range(iterations):
        t0 = time.perf_counter()
        sw.step_device(it)
        t_step += time.perf_counter() - t0
        evaluate_and_reduce()
    val, pos = sw.global_best()
    out = dict(setup_seconds=t_setup, step_seconds=t_step, rank=comm.rank, world=comm.world, particles=(lo, hi), best_value=val, best_position=pos, trace=np.array(trace),
               eval_seconds=t_eval, comm_seconds=t_comm, evaluations=(iterations + 1) * (hi - lo))
    if return_positions:            # the swarm itself stays on the device unless asked for
        sw.fetch()
        out["final_positions"] = sw.positions()
    return out
