#!/usr/bin/env python
"""Timed runs of BASELINE.json configs[2..4] (they are parity-test cases, not bench lines; the numbers go to DESIGN.md).

    python tools/run_configs.py mh   [--chains 4096] [--iterations 50]
    python tools/run_configs.py pso  [--particles 65536] [--iterations 20]
    python tools/run_configs.py ppc16 [--draws 100000] [--chunk 8192]
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 tools/run_configs.py mh ...
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as g

ap = argparse.ArgumentParser()
ap.add_argument("what", choices=["mh", "pso", "ppc16", "ppcq"])
ap.add_argument("--chains", type=int, default=4096)
ap.add_argument("--particles", type=int, default=65536)
ap.add_argument("--iterations", type=int, default=30)
ap.add_argument("--draws", type=int, default=100000)
ap.add_argument("--chunk", type=int, default=8192)
ap.add_argument("--host-swarm", action="store_true", help="pso: keep the swarm on the host (the pre-device-resident path)")
ap.add_argument("--host-staged", action="store_true", help="mh / pso: round 1's path (host sampler resp. per-iteration host round trips, host-staged collectives) "
                                                           "instead of the device-resident callers with the peer-memory exchange")
ap.add_argument("--transport", default=None, choices=[None, "p2p", "nccl"], help="mh / pso: exchange transport of the device-resident callers")
ap.add_argument("--lookahead", type=int, default=1, help="mh: iterations per look-ahead window of the device-resident chains (1: one iteration per launch, 0: from the shard size)")
ap.add_argument("--pageable", action="store_true", help="ppcq: pass the draws in ordinary (pageable) host memory instead of page-locked memory")
ap.add_argument("--ages", type=int, default=4, help="ppcq: 4, or 16 for the synthetic many-age-group variant of BASELINE configs[4]")
a = ap.parse_args()
pkg = g.load_package(); orc = g.load_oracle()
from sepaihrd_b200 import drivers
from sepaihrd_b200.distributed import Comm
from sepaihrd_b200.evaluator import BatchEvaluator
from sepaihrd_b200.problem import TRAJ_FULL

comm = Comm()
dev = comm.local_rank if comm.world > 1 else 0
torch.cuda.set_device(dev)
p = pkg.load_default_problem()
out = dict(what=a.what, world=comm.world)
if a.what in ("mh", "pso") and not (a.host_staged or a.host_swarm):
    # the device-resident callers (csrc/sepaihrd_mh.cu, csrc/sepaihrd_swarm.cu) with the per-iteration exchange on device buffers
    from sepaihrd_b200 import resident
    prob = p.__class__.from_json(dict(p.to_json(), constraint_mode=1)) if a.what == "mh" else p
    with BatchEvaluator(prob, device=dev) as ev:
        ev.eval_batch(np.tile(p.base_params(), (64, 1)))
        if a.what == "mh":
            r = resident.run_mh_resident(ev, p.sigmas, p.base_params(), a.chains, a.iterations, 1234, comm.rank, comm.world, transport=a.transport,
                                         lookahead=None if a.lookahead == 0 else a.lookahead)
            out.update(chains=a.chains, iterations=a.iterations, seconds=r["run_seconds"], setup_seconds=r["setup_seconds"], window_length=r["lookahead"], windows=r["windows"],
                       evals_per_s=a.chains * (a.iterations - 1) / r["run_seconds"], phase_seconds=r["phase_seconds"], transport=r["transport"],
                       accept_rate=float(r["accepts"].mean()), best=float(r["best_trace"][-1]),
                       accept_checksum_local=int(np.packbits(r["accepts"]).astype(np.int64).sum()))
        else:
            r = resident.run_pso_resident(ev, a.particles, a.iterations, 7, initial=p.base_params(), rank=comm.rank, world=comm.world, transport=a.transport)
            out.update(swarm="device, asynchronous", particles=a.particles, iterations=a.iterations, seconds=r["run_seconds"], setup_seconds=r["setup_seconds"],
                       evals_per_s=a.particles * (a.iterations + 1) / r["run_seconds"], phase_seconds=r["phase_seconds"], transport=r["transport"],
                       best_first=float(r["trace"][0]), best_last=float(r["trace"][-1]))
elif a.what == "mh":
    rp = p.__class__.from_json(dict(p.to_json(), constraint_mode=1))
    with BatchEvaluator(rp, device=dev) as ev:
        ev.eval_batch(np.tile(p.base_params(), (64, 1)))
        comm.barrier(); t0 = time.perf_counter()
        r = drivers.run_multichain_mh(ev.eval_batch, p.sigmas, p.lower_bound, p.upper_bound, p.base_params(), a.chains, a.iterations,
                                      seed=1234, comm=comm, record_accepts=True)
        comm.barrier(); dt = time.perf_counter() - t0
    out.update(chains=a.chains, iterations=a.iterations, seconds=dt, iterations_per_s=(a.iterations - 1) / dt,
               evals_per_s=a.chains * (a.iterations - 1) / dt, eval_seconds=r["eval_seconds"], comm_seconds=r["comm_seconds"],
               accept_rate=float(r["accepts"].mean()), best=float(r["best_trace"][-1]),
               accept_checksum=int(np.packbits(r["accepts"]).astype(np.int64).sum()))
elif a.what == "pso":
    with BatchEvaluator(p, device=dev) as ev:
        ev.eval_batch(np.tile(p.base_params(), (64, 1)))
        comm.barrier(); t0 = time.perf_counter()
        r = drivers.run_pso(ev.eval_batch, p.sigmas, p.lower_bound, p.upper_bound, a.particles, a.iterations, seed=7,
                            initial=p.base_params(), comm=comm, device_ctx=None if a.host_swarm else ev.handle)
        comm.barrier(); dt = time.perf_counter() - t0
    out.update(setup_seconds=r.get("setup_seconds"), step_seconds=r.get("step_seconds"))
    out.update(swarm="host" if a.host_swarm else "device", particles=a.particles, iterations=a.iterations, seconds=dt, evals_per_s=a.particles * (a.iterations + 1) / dt,
               eval_seconds=r["eval_seconds"], comm_seconds=r["comm_seconds"], best_first=float(r["trace"][0]), best_last=float(r["trace"][-1]))
elif a.what == "ppcq":
    # posterior-predictive QUANTILES (ResultAggregator): draws -> trajectories -> series -> sort -> quantiles, one C-ABI call
    o = orc.Oracle(p)
    pq = p if a.ages == 4 else p.expand_ages(a.ages // 4)
    oq = orc.Oracle(pq)
    draws = oq.jitter_params(4096, seed=11)
    draws = np.tile(draws, ((a.draws + 4095) // 4096, 1))[:a.draws]
    if not a.pageable:          # page-locked draws, as include/sepaihrd_b200.h recommends for the host-buffer entry points (sepaihrd_alloc_pinned)
        draws = torch.from_numpy(draws).pin_memory().numpy()
    with BatchEvaluator(pq, device=dev) as ev:
        ev.posterior_predictive(draws[:1024], pq.data_initial_state)
        t0 = time.perf_counter()
        q, valid = ev.posterior_predictive(draws, pq.data_initial_state)     # first call of this size: allocates the work buffers
        first = time.perf_counter() - t0
        t0 = time.perf_counter()
        q, valid = ev.posterior_predictive(draws, pq.data_initial_state)     # steady state: the ctx reuses them
        dt = time.perf_counter() - t0
        free_b, total_b = torch.cuda.mem_get_info(dev)
    out.update(draws=a.draws, ages=a.ages, host_memory="pageable" if a.pageable else "pinned", seconds=dt, first_call_seconds=first, draws_per_s=a.draws / dt, valid=valid,
               device_gbytes_in_use=(total_b - free_b) / 1e9, median_deaths_last_day=[float(x) for x in q[2, -1, :, 2]])
else:
    p16 = p.expand_ages(4)
    o16 = orc.Oracle(p16)
    draws = o16.jitter_params(min(a.draws, 4096), seed=11)
    n_local = a.draws // comm.world
    with BatchEvaluator(p16, device=dev) as ev:
        W = p16.state_size; K = p16.n_times
        d_out = torch.empty((a.chunk, K, W), dtype=torch.float64, device=f"cuda:{dev}")
        d_st = torch.empty(a.chunk, dtype=torch.int32, device=f"cuda:{dev}")
        reps = (a.chunk + len(draws) - 1) // len(draws)
        d_par = torch.from_numpy(np.tile(draws, (reps, 1))[:a.chunk]).to(f"cuda:{dev}")
        ev.set_stream(torch.cuda.current_stream().cuda_stream)
        from sepaihrd_b200 import capi
        def launch(nb):
            capi.check(ev._lib.sepaihrd_simulate_batch_device(ev._h, d_par.data_ptr(), nb, d_par.shape[1], TRAJ_FULL, 1, d_out.data_ptr(), d_st.data_ptr()))
        launch(a.chunk); torch.cuda.synchronize()
        comm.barrier(); t0 = time.perf_counter()
        done = 0
        while done < n_local:                      # the trajectories stay on the device (a consumer would read d_out here)
            nb = min(a.chunk, n_local - done)
            launch(nb)
            done += nb
        torch.cuda.synchronize(); comm.barrier(); dt = time.perf_counter() - t0
        # parity on a small subsample against the oracle
        tr, st = ev.simulate_batch(draws[:8])
        ref, _ = o16.simulate_batch(draws[:8])
        rel = float((np.abs(tr - ref) / np.maximum(np.abs(ref), 1.0)).max())
    out.update(draws=a.draws, ages=16, seconds=dt, draws_per_s=a.draws / dt, bytes_per_draw=K * W * 8,
               trajectory_gbytes_per_s=a.draws * K * W * 8 / dt / 1e9, max_rel_vs_oracle=rel)
if comm.rank == 0:
    print(json.dumps(out))
comm.close()
