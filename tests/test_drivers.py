"""Host layer and drivers on the CPU: the C++ samplers (host/optimizers.cpp) stepped from Python, with the CPU oracle
as the objective (tests may use the oracle; the product never does), single process and world-size-2 gloo."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host(pkg, cuda_lib):
    import __graft_entry__ as entry
    entry.build()
    from sepaihrd_b200 import hostlib
    hostlib.load_library()
    return hostlib


@pytest.fixture(scope="module")
def reflect_oracle(problem, orc):
    return orc.Oracle(problem.__class__.from_json(dict(problem.to_json(), constraint_mode=1)))


def test_host_library_exports_every_declared_symbol(host):
    text = open(os.path.join(ROOT, "mathematical-modeling-of-infectious-diseases-v1_b200", "host", "host_capi.h")).read()
    declared = set(re.findall(r"\b(sepaihrd_host_[a-z0-9_]+)\s*\(", text))
    declared -= {"sepaihrd_host_batch_fn"}
    assert declared == set(host.SIGNATURES), declared ^ set(host.SIGNATURES)
    L = host.load_library()
    for name in declared:
        assert hasattr(L, name)


def test_parameter_manager_clamp_and_reflect(host, problem):
    rng = np.random.default_rng(0)
    lo, hi = problem.lower_bound.copy(), problem.upper_bound.copy()
    lo[3] = np.nan; hi[3] = np.nan                                   # a parameter without a bounds entry
    x = problem.base_params() + 3.0 * (hi - np.nan_to_num(lo)).clip(1e-3) * rng.standard_normal(problem.n_params)
    x[3] = -0.25
    pm = host.ParameterManager(problem.sigmas, lo, hi, mode=0)
    c = pm.apply_constraints(x)
    ok = ~np.isnan(lo)
    np.testing.assert_array_equal(c[ok], np.minimum(np.maximum(x[ok], lo[ok]), hi[ok]))
    assert c[3] == 0.0                                                # max(0, v)
    pm.set_mode(1)
    r = pm.apply_constraints(x)
    w = hi[ok] - lo[ok]
    y = np.fmod(x[ok] - lo[ok], 2 * w); y = np.where(y < 0, y + 2 * w, y)
    np.testing.assert_array_equal(r[ok], np.where(y <= w, lo[ok] + y, hi[ok] - (y - w)))
    assert r[3] == 0.25                                               # |v|
    assert np.all((r[ok] >= lo[ok]) & (r[ok] <= hi[ok]))


def _mh(host, problem, evaluate, n_chains, iterations, seed, offset=0, total=None):
    pm = host.ParameterManager(problem.sigmas, problem.lower_bound, problem.upper_bound, mode=1)
    mh = host.MultiChainMH(pm, dict(mcmc_iterations=iterations, burn_in=iterations, n_chains=n_chains, chain_offset=offset, seed=seed, store_samples=0))
    x0 = problem.base_params()
    mh.begin(x0, np.full(n_chains, evaluate(x0[None])[0]))
    acc, props = [], []
    while not mh.done:
        p = mh.propose(); props.append(p)
        acc.append(mh.accept(evaluate(p)))
    return mh, np.array(acc), np.array(props)


def test_multichain_mh_is_repeatable_and_chain_streams_are_global(host, problem, reflect_oracle):
    ev = lambda x: reflect_oracle.eval_batch(x)[0]
    mh_a, acc_a, prop_a = _mh(host, problem, ev, 6, 8, seed=5)
    mh_b, acc_b, prop_b = _mh(host, problem, ev, 6, 8, seed=5)
    np.testing.assert_array_equal(acc_a, acc_b)
    np.testing.assert_array_equal(prop_a, prop_b)
    # chains 3..5 run as their own shard visit the same states (global chain index seeds the stream)
    mh_c, acc_c, prop_c = _mh(host, problem, ev, 3, 8, seed=5, offset=3)
    np.testing.assert_array_equal(acc_c, acc_a[:, 3:])
    np.testing.assert_array_equal(prop_c, prop_a[:, 3:])
    # a different seed gives different proposals; proposals respect the bounds (reflection)
    _, _, prop_d = _mh(host, problem, ev, 6, 3, seed=6)
    assert not np.array_equal(prop_d[0], prop_a[0])
    assert np.all(prop_a >= problem.lower_bound) and np.all(prop_a <= problem.upper_bound)
    x, lp, scale, n_acc = mh_a.state()
    np.testing.assert_array_equal(n_acc, acc_a.sum(axis=0))
    np.testing.assert_allclose(lp, ev(x), rtol=0, atol=0)             # the carried log-posterior is the one of the carried state
    assert np.all(scale > 0)


def test_mh_follows_the_reference_accept_rule(host, problem):
    """Synthetic objective: uphill proposals are always taken, hopeless ones never; NaN/inf count as -1e18."""
    calls = {"n": 0}

    def ev(x):
        calls["n"] += 1
        return np.full(len(x), 10.0 * calls["n"])                     # strictly increasing: every proposal is uphill
    _, acc, _ = _mh(host, problem, ev, 4, 6, seed=1)
    assert acc.all()
    calls["n"] = 0

    def ev_down(x):
        calls["n"] += 1
        return np.full(len(x), 0.0 if calls["n"] == 1 else np.nan)    # initial 0, then NaN -> -1e18 -> never accepted
    _, acc, _ = _mh(host, problem, ev_down, 4, 6, seed=1)
    assert not acc.any()


def test_adaptive_covariance_path_runs(host, problem):
    """iterations > burn_in: rank-1 updates every iteration and a full recomputation + Cholesky every adaptation period."""
    target = problem.base_params()
    sc = np.maximum(problem.sigmas, 1e-9)
    ev = lambda x: -0.5 * (((x - target) / sc) ** 2).sum(axis=1)
    pm = host.ParameterManager(problem.sigmas, problem.lower_bound, problem.upper_bound, mode=1)
    mh = host.MultiChainMH(pm, dict(mcmc_iterations=260, burn_in=100, adaptation_period=50, n_chains=2, seed=3))
    mh.begin(target, ev(target[None]).repeat(2))
    while not mh.done:
        mh.accept(ev(mh.propose()))
    x, lp, scale, n_acc = mh.state()
    assert np.isfinite(lp).all() and (n_acc > 0).all() and (n_acc < 259).all()
    assert mh.best()[1] <= 0.0


def test_pso_finds_the_optimum_of_a_concave_objective_and_shards_agree(host, problem):
    target = 0.5 * (problem.lower_bound + problem.upper_bound)
    width = (problem.upper_bound - problem.lower_bound)
    ev = lambda x: -(((x - target) / width) ** 2).sum(axis=1)
    pm = host.ParameterManager(problem.sigmas, problem.lower_bound, problem.upper_bound, mode=0)

    def run(lo, hi, hook=None):
        sw = host.Swarm(pm, dict(iterations=30, swarm_size=24, particle_offset=lo, local_count=hi - lo, seed=9))
        sw.begin(None)
        return sw

    # single process
    sw = run(0, 24)
    trace = []
    v, i, pos = sw.tell(ev(sw.positions())); sw.set_global_best(v, pos)
    first = sw.global_best()[0]
    for it in range(30):
        sw.step(it)
        v, i, pos = sw.tell(ev(sw.positions())); sw.set_global_best(v, pos)
        trace.append(sw.global_best()[0])
    assert all(b >= a for a, b in zip(trace, trace[1:])) and trace[-1] > first
    assert np.all(sw.positions() >= problem.lower_bound) and np.all(sw.positions() <= problem.upper_bound)
    # two shards stepped side by side with a manual global-best exchange reproduce the single-process trace
    a, b = run(0, 11), run(11, 24)
    trace2 = []

    def exchange():
        va, ia, pa = a.tell(ev(a.positions())); vb, ib, pb = b.tell(ev(b.positions()))
        gv, gp = (va, pa) if va >= vb else (vb, pb)
        a.set_global_best(gv, gp); b.set_global_best(gv, gp)
        return a.global_best()[0]
    exchange()
    for it in range(30):
        a.step(it); b.step(it)
        trace2.append(exchange())
    np.testing.assert_array_equal(trace, trace2)


def test_whole_run_optimizers_improve_the_oracle_likelihood(host, problem, oracle):
    ev = lambda x: oracle.eval_batch(x)[0]
    pm = host.ParameterManager(problem.sigmas, problem.lower_bound, problem.upper_bound, mode=0)
    x0 = problem.base_params()
    f0 = ev(x0[None])[0]
    best, val, nev = host.optimize("hill", pm, dict(iterations=3, cloud_size=16, seed=4), ev, x0)
    assert val >= f0 and nev >= 1 + 3 * 16
    np.testing.assert_allclose(ev(best[None])[0], val, rtol=1e-13)
    best, val, nev = host.optimize("pso", pm, dict(iterations=3, swarm_size=12, seed=4, **host.BASIC_SWARM), ev, x0)
    assert val >= f0 and nev == 12 * 4                                  # particle 0 starts at the initial point
    # the class defaults are the reference's: ADAPTIVE variant, opposition-based initialisation (the swarm is scored twice),
    # adaptive coefficients, elitist learning after iteration 0 (3 trials)
    best, val, nev = host.optimize("pso", pm, dict(iterations=3, swarm_size=12, seed=4), ev, x0)
    assert val >= f0 and nev == 12 * 2 + 3 * 12 + 3
    pm.set_mode(1)
    best, val, nev = host.optimize("mh", pm, dict(mcmc_iterations=6, burn_in=6, n_chains=3, seed=4, lookahead=1), ev, x0)
    assert val >= f0 and nev == 1 + 5 * 3
    best2, val2, nev2 = host.optimize("mh", pm, dict(mcmc_iterations=6, burn_in=6, n_chains=3, seed=4), ev, x0)      # look-ahead (the default)
    assert val2 == val and np.array_equal(best2, best) and nev2 > nev
    with pytest.raises(host.HostError):
        host.optimize("pso", pm, dict(iterations=1, swarm_size=4, variant=5), ev, x0)      # "variant must be between 0 and 4"


def _spawn(what, tmp_path, world, port):
    out = str(tmp_path / what)
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_dist_worker.py"), what, out], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    for p in procs:
        o, _ = p.communicate(timeout=600)
        assert p.returncode == 0, o.decode()[-3000:]
    return [np.load(f"{out}.rank{r}.npz") for r in range(world)]


@pytest.mark.timeout(900)
def test_world_size_two_gloo_matches_single_process(tmp_path, host, problem):
    """Sharding over ranks (gloo, world 2) visits exactly the states of the single-process run, for both drivers."""
    one = _spawn("mh", tmp_path / "w1", 1, 29611) if (tmp_path / "w1").mkdir() is None else None
    two = _spawn("mh", tmp_path / "w2", 2, 29612) if (tmp_path / "w2").mkdir() is None else None
    ref = one[0]
    assert (int(two[0]["lo"]), int(two[0]["hi"]), int(two[1]["lo"]), int(two[1]["hi"])) == (0, 3, 3, 6)
    np.testing.assert_array_equal(np.concatenate([two[0]["accepts"], two[1]["accepts"]], axis=1), ref["accepts"])
    np.testing.assert_array_equal(np.concatenate([two[0]["x"], two[1]["x"]]), ref["x"])
    for r in two:
        np.testing.assert_array_equal(r["all_logpost"], ref["logpost"])    # the per-iteration gather gives every rank every chain
    one = _spawn("pso", tmp_path / "w1", 1, 29613)
    two = _spawn("pso", tmp_path / "w2", 2, 29614)
    for r in two:
        np.testing.assert_array_equal(r["trace"], one[0]["trace"])
        np.testing.assert_array_equal(r["best_position"], one[0]["best_position"])


def test_two_phase_model_calibrator_on_a_synthetic_posterior(host, problem):
    """ModelCalibrator::calibrate on the CPU with a synthetic Gaussian log-posterior: phase 1 (PSO or hill climbing, clamp mode)
    hands its covariance to phase 2 (Metropolis-Hastings, reflect mode, eigenvalues floored at (0.1 sigma)^2, x4 inflation), and
    every stored MCMC sample is re-scored in one batch."""
    mid = 0.5 * (problem.lower_bound + problem.upper_bound)
    width = problem.upper_bound - problem.lower_bound
    target = mid + 0.1 * width
    ev = lambda x: -0.5 * (((x - target) / (0.05 * width)) ** 2).sum(axis=1)
    pm = host.ParameterManager(problem.sigmas, problem.lower_bound, problem.upper_bound, mode=0)
    f0 = ev(mid[None])[0]
    for phase1, s1 in (("pso", dict(iterations=25, swarm_size=40, seed=2)), ("hill", dict(iterations=15, cloud_size=32, seed=2))):
        best, val, n_samples, p1 = host.calibrate(phase1, pm, s1, dict(mcmc_iterations=40, burn_in=40, n_chains=4, seed=3, thinning=2), ev, mid)
        assert p1 > f0 and val >= p1                      # phase 1 improves on the start, the overall best is at least that
        assert n_samples == 4 * (1 + 39 // 2)             # initial state + every 2nd iteration, per chain
        np.testing.assert_allclose(ev(best[None])[0], val, rtol=1e-12)
        assert np.all(best >= problem.lower_bound) and np.all(best <= problem.upper_bound)


def test_metropolis_hastings_writes_the_reference_trace_files(host, problem, tmp_path):
    """posterior_trace_checkpoint.csv / posterior_trace_final.csv / posterior_trace.csv of MetropolisHastingsSampler::optimize
    (.cpp:380-382, 399-409, 414-469): iter, log_posterior and one column per parameter, %.6e, thinning applied."""
    target = problem.base_params()
    sc = np.maximum(problem.sigmas, 1e-9)
    ev = lambda x: -0.5 * (((x - target) / sc) ** 2).sum(axis=1)
    pm = host.ParameterManager(problem.sigmas, problem.lower_bound, problem.upper_bound, mode=1)
    st = dict(mcmc_iterations=60, burn_in=10, adaptation_period=20, n_chains=1, report_interval=20, thinning=2, seed=1)
    out = tmp_path / "traces"
    host.set_trace_directory(str(out))
    try:
        best, val, nev = host.optimize("mh", pm, st, ev, target)
        names = sorted(p.name for p in out.iterdir())
        assert names == ["posterior_trace.csv", "posterior_trace_checkpoint.csv", "posterior_trace_final.csv"]
        final = (out / "posterior_trace_final.csv").read_text().splitlines()
        assert final[0].split(",")[:2] == ["iter", "log_posterior"] and len(final[0].split(",")) == 2 + problem.n_params
        assert len(final) == 1 + 1 + 29                           # header, the initial sample, t = 2, 4, ..., 58
        assert (out / "posterior_trace.csv").read_text().splitlines() == final
        rows = np.array([[float(v) for v in ln.split(",")] for ln in final[1:]])
        np.testing.assert_array_equal(rows[:, 0], np.arange(30))
        np.testing.assert_allclose(rows[:, 1], ev(rows[:, 2:]), rtol=2e-5, atol=1e-6)      # six significant digits
        assert rows[:, 1].max() == pytest.approx(val, rel=1e-6)
        assert all(len(v.split("e")[0]) <= 9 for v in final[1].split(",")[1:])              # d.dddddde+xx
        ckpt = (out / "posterior_trace_checkpoint.csv").read_text().splitlines()
        assert ckpt[0] == final[0] and 1 < len(ckpt) <= len(final)                           # written at t + 1 = 20, 40, 60
        # write_trace 0 / write_checkpoints 0, and no directory known outside a reference-style tree: nothing is written
        for p in out.iterdir():
            p.unlink()
        host.optimize("mh", pm, dict(st, write_trace=0, write_checkpoints=0), ev, target)
        assert list(out.iterdir()) == []
        host.optimize("mh", pm, dict(st, write_trace=0), ev, target)
        assert sorted(p.name for p in out.iterdir()) == ["posterior_trace_checkpoint.csv", "posterior_trace_final.csv"]
    finally:
        host.set_trace_directory(None)
    before = set(os.listdir(ROOT))
    host.optimize("mh", pm, st, ev, target)
    assert set(os.listdir(ROOT)) == before and not os.path.exists(os.path.join(os.getcwd(), "data", "mcmc_samples"))


def test_master_generator_draws_for_the_resident_swarm_are_std_mt19937():
    """resident.std_mt19937_raw (numpy's MT19937 with init_genrand seeding) against the Python restatement of std::mt19937 the
    sampler tests are built on: the device-resident swarm takes its per-particle seeds from these draws."""
    import __graft_entry__ as entry
    entry.load_package()
    from sepaihrd_b200 import resident
    from test_pso_variants import StdMt19937
    for seed in (0, 1, 7, 5489, 2 ** 32 - 1):
        g = StdMt19937(seed)
        want = np.array([g.raw() for _ in range(1500)], dtype=np.uint32)
        np.testing.assert_array_equal(resident.std_mt19937_raw(seed, 1500), want)
    np.testing.assert_array_equal(resident.initial_cholesky([0.5, 0.0, 2.0]),
                                  np.diag(np.sqrt(np.array([0.25, 1e-6, 4.0]) * (2.38 * 2.38 / 3.0) + 1e-6)))
