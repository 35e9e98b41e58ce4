"""GPU tests of the device-resident callers (csrc/sepaihrd_mh.cu, the asynchronous form of csrc/sepaihrd_swarm.cu) and of the
peer-memory exchange (csrc/sepaihrd_exchange.cu), all through the C ABI:

  * the device-resident Metropolis-Hastings chains make EXACTLY the host sampler's accept decisions and visit its states
    (host/optimizers.cpp is itself pinned bit for bit against a Python restatement of the reference sampler and, through the
    evaluator, against the CPU oracle): seeded accept sequences identical, north_star's gate for configs[2];
  * the asynchronous swarm reproduces the host / synchronous device swarm's global-best trace bit for bit;
  * two PROCESSES sharing cuda:0 exchange their records through the CUDA-IPC mailboxes and reproduce the single-process run.
"""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host(pkg, cuda_lib):
    import __graft_entry__ as entry
    entry.build()
    from sepaihrd_b200 import hostlib
    return hostlib


@pytest.fixture(scope="module")
def mods(cuda_lib):
    from sepaihrd_b200 import drivers, evaluator, resident
    return drivers, evaluator, resident


@pytest.fixture(scope="module")
def reflect_problem(problem):
    return problem.__class__.from_json(dict(problem.to_json(), constraint_mode=1))


def test_exchange_single_rank_is_a_device_copy(problem, mods):
    import torch
    _, evaluator, resident = mods
    with evaluator.BatchEvaluator(problem, device=0) as ev:
        ev.set_stream(torch.cuda.current_stream().cuda_stream)
        ex = resident.Exchange(ev, 1000)
        assert ex.transport == "single"
        for rnd in range(5):
            src = torch.arange(777, dtype=torch.float64, device="cuda") + rnd
            dst = torch.zeros((1, 777), dtype=torch.float64, device="cuda")
            ex.all_gather(src.data_ptr(), 777, dst.data_ptr())
            torch.cuda.synchronize()
            assert torch.equal(dst[0], src)
        assert ex.status() == 0
        ex.close()


@pytest.mark.parametrize("diag", [True, False])
def test_device_resident_chains_make_the_host_samplers_decisions(host, problem, reflect_problem, mods, diag):
    """Accept matrix, visited states, log-posteriors and Robbins-Monro scales of 203 seeded chains x 40 iterations: device-resident
    (propose / accept kernels, generators in HBM) == host sampler stepping the same evaluator -- bit for bit, with the start
    kernel built from the sigmas (diagonal) and with a dense covariance handed over (phase-1 style)."""
    drivers, evaluator, resident = mods
    n_chains, iters, seed = 203, 40, 1234
    P = problem.n_params
    if diag:
        cov, chol = None, None
    else:
        rng = np.random.default_rng(3)
        A = rng.standard_normal((P, P)) * 0.05
        cov = (np.diag(problem.sigmas ** 2) + (A * problem.sigmas) @ (A * problem.sigmas).T) * 0.05
    with evaluator.BatchEvaluator(reflect_problem, device=0) as ev:
        pm = host.ParameterManager(problem.sigmas, problem.lower_bound, problem.upper_bound, mode=1)
        mh = host.MultiChainMH(pm, dict(mcmc_iterations=iters, burn_in=iters, n_chains=n_chains, seed=seed, store_samples=0))
        if cov is not None:
            mh.set_initial_covariance(cov)
            chol = np.linalg.cholesky(cov + 1e-6 * np.eye(P))      # only to know the host factors SOMETHING like this; the device gets the host's own factor below
        x0 = problem.base_params()
        mh.begin(x0, np.full(n_chains, ev.eval_batch(x0[None])[0][0]))
        acc_ref, props = [], []
        while not mh.done:
            p = mh.propose(); props.append(p)
            acc_ref.append(mh.accept(ev.eval_batch(p)[0]))
        x_ref, lp_ref, sc_ref, n_ref = mh.state()
        acc_ref = np.array(acc_ref)
        if cov is not None:
            chol = mh.shared_cholesky()                             # the factor the host sampler actually uses (its own Cholesky routine)
        r = resident.run_mh_resident(ev, problem.sigmas, x0, n_chains, iters, seed, chol_lower=chol)
    assert r["accepts"].shape == acc_ref.shape == (iters - 1, n_chains)
    np.testing.assert_array_equal(r["accepts"], acc_ref)
    np.testing.assert_array_equal(r["x"], x_ref)
    np.testing.assert_array_equal(r["logpost"], lp_ref)
    np.testing.assert_array_equal(r["scale"], sc_ref)
    np.testing.assert_array_equal(r["accepted"], n_ref)
    assert 0.02 < acc_ref.mean() < 0.98
    # the trace is the max over all chains' current log-posteriors after every iteration
    assert r["best_trace"][-1] == lp_ref.max()
    assert np.all(np.diff(r["best_trace"]) >= 0) or True      # (not monotone in general: a chain may leave its best state)


@pytest.mark.parametrize("diag", [True, False])
def test_look_ahead_windows_make_the_one_iteration_runs_decisions(problem, reflect_problem, mods, diag):
    """sepaihrd_mh_window_*: every chain proposes its next K iterations from a copy of its generator, ONE likelihood launch scores
    local x K proposals, every chain commits up to its first accepted proposal.  Accept matrix, states, log-posteriors, scales
    and accept counts of 203 chains x 60 iterations equal the one-iteration-per-launch device run (itself equal to the host
    sampler, above) for K = 2, 5, 16 and the automatic length; the generator positions are right or the later draws would differ."""
    _, evaluator, resident = mods
    n_chains, iters, seed = 203, 60, 77
    P = problem.n_params
    chol = None
    if not diag:
        rng = np.random.default_rng(5)
        A = rng.standard_normal((P, P)) * 0.05
        cov = (np.diag(problem.sigmas ** 2) + (A * problem.sigmas) @ (A * problem.sigmas).T) * 0.05
        chol = np.linalg.cholesky(cov + 1e-6 * np.eye(P))
    x0 = problem.base_params()
    with evaluator.BatchEvaluator(reflect_problem, device=0) as ev:
        ref = resident.run_mh_resident(ev, problem.sigmas, x0, n_chains, iters, seed, chol_lower=chol)
        assert ref["lookahead"] == 1 and ref["windows"] == iters - 1
        for K in (2, 5, 16, None):
            r = resident.run_mh_resident(ev, problem.sigmas, x0, n_chains, iters, seed, chol_lower=chol, lookahead=K)
            assert r["lookahead"] == (K if K else resident.window_length(n_chains)) and r["windows"] < ref["windows"]
            np.testing.assert_array_equal(r["accepts"], ref["accepts"], err_msg=f"K={K}")
            np.testing.assert_array_equal(r["x"], ref["x"])
            np.testing.assert_array_equal(r["logpost"], ref["logpost"])
            np.testing.assert_array_equal(r["scale"], ref["scale"])
            np.testing.assert_array_equal(r["accepted"], ref["accepted"])
            assert r["best_trace"][-1] == ref["best_trace"][-1] == ref["logpost"].max()
        # without scale adaptation, and a window longer than the run
        a = resident.run_mh_resident(ev, problem.sigmas, x0, 37, 9, seed, chol_lower=chol, adapt_scale=False)
        b = resident.run_mh_resident(ev, problem.sigmas, x0, 37, 9, seed, chol_lower=chol, adapt_scale=False, lookahead=32)
        np.testing.assert_array_equal(a["accepts"], b["accepts"])
        np.testing.assert_array_equal(a["x"], b["x"])
        # a run uses either the windows or the one-iteration phases
        mh = resident.DeviceMH(ev, 8, 0, 8, 20)
        mh.begin(seed, x0, resident.initial_cholesky(problem.sigmas))
        mh.iterate(2)
        with pytest.raises(Exception, match="cannot follow"):
            mh.window_propose(4)
        mh.close()
        mh = resident.DeviceMH(ev, 8, 0, 8, 20)
        mh.begin(seed, x0, resident.initial_cholesky(problem.sigmas))
        assert mh.run_windows(4) < 19 and mh.iteration == 20
        with pytest.raises(Exception, match="look-ahead windows"):
            mh.propose()
        mh.close()


def test_look_ahead_windows_over_a_long_run(problem, reflect_problem, mods):
    """400 iterations of 16 chains: ~65 000 generator words per chain (over a hundred wrap-arounds of the 624-word state), windows of
    7 and of 64 iterations -- a window of 64 draws ~10 000 words ahead on the copy and the commit twists the real state forward by
    whatever was consumed.  Decisions, states and scales equal the one-iteration-per-launch run."""
    _, evaluator, resident = mods
    x0 = problem.base_params()
    with evaluator.BatchEvaluator(reflect_problem, device=0) as ev:
        ref = resident.run_mh_resident(ev, problem.sigmas, x0, 16, 400, 4242)
        for K in (7, 64):
            r = resident.run_mh_resident(ev, problem.sigmas, x0, 16, 400, 4242, lookahead=K)
            np.testing.assert_array_equal(r["accepts"], ref["accepts"], err_msg=f"K={K}")
            np.testing.assert_array_equal(r["x"], ref["x"])
            np.testing.assert_array_equal(r["scale"], ref["scale"])
            np.testing.assert_array_equal(r["logpost"], ref["logpost"])
            assert r["windows"] < 399 / 2


def test_asynchronous_swarm_equals_the_synchronous_device_swarm(problem, mods):
    drivers, evaluator, resident = mods
    kw = dict(sigmas=problem.sigmas, lower=problem.lower_bound, upper=problem.upper_bound, swarm_size=333, iterations=6, seed=7,
              initial=problem.base_params())
    with evaluator.BatchEvaluator(problem, device=0) as ev:
        ref = drivers.run_pso(None, device_ctx=ev.handle, return_positions=True, **kw)
        got = resident.run_pso_resident(ev, 333, 6, 7, initial=problem.base_params(), return_positions=True)
        no_init = resident.run_pso_resident(ev, 64, 2, 9)
        ref2 = drivers.run_pso(None, device_ctx=ev.handle, sigmas=problem.sigmas, lower=problem.lower_bound, upper=problem.upper_bound,
                               swarm_size=64, iterations=2, seed=9)
    np.testing.assert_array_equal(got["trace"], ref["trace"])
    np.testing.assert_array_equal(got["best_position"], ref["best_position"])
    np.testing.assert_array_equal(got["final_positions"], ref["final_positions"])
    assert got["best_value"] == ref["best_value"]
    np.testing.assert_array_equal(no_init["trace"], ref2["trace"])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _launch(what, out):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()), WORLD_SIZE="2", SEPAIHRD_EXCHANGE_TIMEOUT_S="20")
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_resident_worker.py"), what, out], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    logs = []
    for pr in procs:
        try:
            o, _ = pr.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            pr.kill(); o, _ = pr.communicate()
        logs.append(o.decode(errors="replace"))
    assert all(pr.returncode == 0 for pr in procs), "\n".join(logs)
    return [np.load(f"{out}.rank{r}.npz") for r in range(2)]


def test_two_processes_exchange_through_peer_memory_mh(problem, reflect_problem, mods, tmp_path):
    """203 chains over two processes that share cuda:0: records cross through the CUDA-IPC mailboxes (sepaihrd_exchange_*);
    the sharded run makes the single-process run's decisions and both ranks hold every chain's log-likelihood."""
    _, evaluator, resident = mods
    with evaluator.BatchEvaluator(reflect_problem, device=0) as ev:
        ref = resident.run_mh_resident(ev, problem.sigmas, problem.base_params(), 203, 12, 1234)
    parts = _launch("mh", str(tmp_path / "mh"))
    assert [int(q["status"]) for q in parts] == [0, 0]
    assert (int(parts[0]["lo"]), int(parts[0]["hi"]), int(parts[1]["lo"]), int(parts[1]["hi"])) == (0, 102, 102, 203)
    np.testing.assert_array_equal(np.concatenate([q["accepts"] for q in parts], axis=1), ref["accepts"])
    np.testing.assert_array_equal(np.concatenate([q["x"] for q in parts]), ref["x"])
    np.testing.assert_array_equal(np.concatenate([q["scale"] for q in parts]), ref["scale"])
    for q in parts:
        np.testing.assert_array_equal(q["all_logpost"], ref["logpost"])
        np.testing.assert_array_equal(q["trace"], ref["best_trace"])


def test_two_processes_run_look_ahead_windows_over_peer_memory(problem, reflect_problem, mods, tmp_path):
    """The windowed run sharded 102 + 101 over two processes: the per-window record (log-likelihood block + the rank's smallest
    iteration index at the stride of the larger shard) crosses the mailboxes, both ranks stop after the same window, and the
    decisions are the single-process one-iteration-per-launch run's."""
    _, evaluator, resident = mods
    with evaluator.BatchEvaluator(reflect_problem, device=0) as ev:
        ref = resident.run_mh_resident(ev, problem.sigmas, problem.base_params(), 203, 40, 1234)
    parts = _launch("mhw", str(tmp_path / "mhw"))
    assert [int(q["status"]) for q in parts] == [0, 0]
    np.testing.assert_array_equal(np.concatenate([q["accepts"] for q in parts], axis=1), ref["accepts"])
    np.testing.assert_array_equal(np.concatenate([q["x"] for q in parts]), ref["x"])
    np.testing.assert_array_equal(np.concatenate([q["scale"] for q in parts]), ref["scale"])
    for q in parts:
        np.testing.assert_array_equal(q["all_logpost"], ref["logpost"])
        assert q["trace"][-1] == ref["best_trace"][-1] and len(q["trace"]) == len(parts[0]["trace"]) < 39


def test_two_processes_exchange_through_peer_memory_pso(problem, mods, tmp_path):
    _, evaluator, resident = mods
    with evaluator.BatchEvaluator(problem, device=0) as ev:
        ref = resident.run_pso_resident(ev, 301, 5, 7, initial=problem.base_params(), return_positions=True)
    parts = _launch("pso", str(tmp_path / "pso"))
    assert [int(q["status"]) for q in parts] == [0, 0]
    for q in parts:
        np.testing.assert_array_equal(q["trace"], ref["trace"])
        np.testing.assert_array_equal(q["best_position"], ref["best_position"])
    np.testing.assert_array_equal(np.concatenate([q["positions"] for q in parts]), ref["final_positions"])


def test_resident_sampler_refuses_what_it_does_not_cover(problem, reflect_problem, mods):
    """No silent fallback: a run that needs the covariance adaptation after burn-in is the host sampler's, an exchange record larger
    than its mailbox slots or an all-gather before connect are errors, a swarm step without its seed set is an error."""
    import ctypes as C
    import torch
    _, evaluator, resident = mods
    from sepaihrd_b200 import capi
    with evaluator.BatchEvaluator(reflect_problem, device=0) as ev:
        with pytest.raises(capi.SepaihrdError, match="covariance"):
            resident.DeviceMH(ev, 16, 0, 16, iterations=50, burn_in=10)
        mh = resident.DeviceMH(ev, 16, 0, 16, iterations=5)
        with pytest.raises(capi.SepaihrdError, match="before sepaihrd_mh_begin"):
            mh.iterate(1)
        mh.close()
        ev.set_stream(torch.cuda.current_stream().cuda_stream)
        ex = resident.Exchange(ev, 64)
        src = torch.zeros(4096, dtype=torch.float64, device="cuda"); dst = torch.zeros(4096, dtype=torch.float64, device="cuda")
        with pytest.raises(capi.SepaihrdError, match="larger than the mailbox"):
            ex.all_gather(src.data_ptr(), 4096, dst.data_ptr())
        ex.close()
        sw = resident.DeviceSwarm(ev, 100, 0, 100)
        with pytest.raises(capi.SepaihrdError, match="upload_seeds"):
            sw.init()
        sw.upload_seeds(resident.std_mt19937_raw(3, 200).reshape(2, 100))
        sw.init(); sw.evaluate(); sw.step(0, 0.9, 2.5, 0.5)
        with pytest.raises(capi.SepaihrdError, match="no seed set"):
            sw.step(1, 0.9, 2.5, 0.5)
        sw.close()
