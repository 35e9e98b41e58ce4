"""Look-ahead Metropolis-Hastings (host/optimizers.cpp `runLookahead`): the reference's shipped configuration is ONE chain of
sequential iterations (data/configuration/mcmc_settings.txt, MetropolisHastingsSampler.cpp:283-384), one objective evaluation
per iteration.  With setting ``lookahead`` != 1 the sampler evaluates the proposals of the next K iterations -- all known in
advance as long as the chain keeps rejecting -- as ONE batch and commits up to the first accepted one.  The chain must be the
sequential chain bit for bit: every visited state, every log-posterior, the generator (a misaligned draw would change every
later state), the scale, the adapted covariance, and the trace / checkpoint files; only the shape of the evaluation batches
differs."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def host(pkg, cuda_lib):
    import __graft_entry__ as entry
    entry.build()
    from sepaihrd_b200 import hostlib
    hostlib.load_library()
    return hostlib


def _objective(problem, kind):
    target = problem.base_params()
    sc = np.maximum(problem.sigmas, 1e-9)
    if kind == "gauss":                      # smooth: acceptance near the target rate once the scale has adapted
        return lambda x: -0.5 * (((x - target) / (3.0 * sc)) ** 2).sum(axis=1)
    if kind == "steep":                      # almost everything is rejected: long windows, the emergency / fast-shrink branches
        return lambda x: -0.5 * (((x - target) / (1e-3 * sc)) ** 2).sum(axis=1)
    if kind == "plateau":                    # ties: log_ratio == 0 accepts WITHOUT drawing a uniform (.cpp:323-329)
        return lambda x: -np.floor(np.abs((x - target) / sc).sum(axis=1))
    if kind == "nan":                        # non-finite values go through safeEvaluate's -1e18 (.cpp:65-74)
        def f(x):
            v = -0.5 * (((x - target) / (3.0 * sc)) ** 2).sum(axis=1)
            v[(np.abs(x[:, 0] - target[0]) / sc[0]) > 1.0] = np.nan
            return v
        return f
    raise ValueError(kind)


def _run(host, problem, tmp_path, tag, kind, settings):
    calls = []
    inner = _objective(problem, kind)

    def ev(x):
        calls.append(x.copy())
        return inner(x)

    pm = host.ParameterManager(problem.sigmas, problem.lower_bound, problem.upper_bound, mode=1)
    out = tmp_path / tag
    host.set_trace_directory(str(out))
    try:
        best, val, nev = host.optimize("mh", pm, settings, ev, problem.base_params())
    finally:
        host.set_trace_directory(None)
    files = {p.name: p.read_text() for p in out.iterdir()} if out.exists() else {}
    return dict(best=best, val=val, nev=nev, files=files, calls=calls)


@pytest.mark.parametrize("kind", ["gauss", "steep", "plateau", "nan"])
@pytest.mark.parametrize("lookahead", [0, 2, 7, 64])
def test_lookahead_chain_is_the_sequential_chain(host, problem, tmp_path, kind, lookahead):
    # 330 iterations: through the burn-in (rank-1 updates from t = 121), three refactorisations of the proposal kernel
    # (t = 150, 200, ... : windows must stop in front of them), checkpoints every 40, thinning 1 so that every state is written
    st = dict(mcmc_iterations=330, burn_in=120, adaptation_period=50, n_chains=1, report_interval=40, thinning=1, seed=11)
    seq = _run(host, problem, tmp_path, "seq", kind, dict(st, lookahead=1))
    la = _run(host, problem, tmp_path, f"la{lookahead}", kind, dict(st, lookahead=lookahead))
    assert all(len(c) == 1 for c in seq["calls"]) and seq["nev"] == 330             # 1 initial + 329 proposals, one per call
    assert sorted(la["files"]) == sorted(seq["files"]) == ["posterior_trace.csv", "posterior_trace_checkpoint.csv", "posterior_trace_final.csv"]
    for name in seq["files"]:
        assert la["files"][name] == seq["files"][name], f"{name} differs (lookahead {lookahead}, {kind})"
    np.testing.assert_array_equal(la["best"], seq["best"])
    assert la["val"] == seq["val"]
    # every proposal the sequential run evaluated appears, bit for bit and in order, among the look-ahead evaluations
    # (the first row of every window is the sequential proposal of that iteration)
    flat = np.concatenate(la["calls"])
    want = np.concatenate(seq["calls"])
    pos = 0
    for row in want:
        while pos < len(flat) and not np.array_equal(flat[pos], row):
            pos += 1
        assert pos < len(flat), "a sequential proposal was never evaluated by the look-ahead run"
        pos += 1
    assert len(la["calls"]) < len(seq["calls"])                                       # fewer launches
    if lookahead > 1:
        assert max(len(c) for c in la["calls"]) <= lookahead


def test_lookahead_cuts_the_number_of_launches(host, problem, tmp_path):
    """On a smooth target with the scale adapted to ~23 % acceptance a window commits ~4 iterations: the number of objective
    calls (device launches behind the drop-in) falls accordingly, at the price of evaluations thrown away."""
    st = dict(mcmc_iterations=3000, burn_in=3000, n_chains=1, seed=5, write_trace=0, write_checkpoints=0)
    la = _run(host, problem, tmp_path, "la", "gauss", dict(st, lookahead=0))
    launches = len(la["calls"]) - 1
    assert 2999 / launches > 1.5, (launches, la["nev"])            # (a device launch costs ~0.6 ms: ~4.5 iterations per launch there)
    assert la["nev"] < 25 * launches


@pytest.mark.parametrize("n_chains", [3, 12])
@pytest.mark.parametrize("kind", ["gauss", "plateau"])
def test_multichain_lookahead_equals_the_lockstep_run(host, problem, tmp_path, n_chains, kind):
    """Several chains: every chain looks ahead on its own (its own iteration counter; the chains run apart between launches).
    Samples of all chains, checkpoints (every chain cut at the reporting iteration) and best value equal the lockstep run's."""
    st = dict(mcmc_iterations=230, burn_in=90, adaptation_period=40, n_chains=n_chains, report_interval=50, thinning=3, seed=21)
    lock = _run(host, problem, tmp_path, "lock", kind, dict(st, lookahead=1))
    assert lock["nev"] == 1 + n_chains * 229 and all(len(c) == n_chains for c in lock["calls"][1:])
    for la in (0, 5):
        r = _run(host, problem, tmp_path, f"la{la}", kind, dict(st, lookahead=la))
        assert sorted(r["files"]) == sorted(lock["files"]) == ["posterior_trace.csv", "posterior_trace_checkpoint.csv", "posterior_trace_final.csv"]
        for name in lock["files"]:
            assert r["files"][name] == lock["files"][name], f"{name} differs ({n_chains} chains, lookahead {la}, {kind})"
        np.testing.assert_array_equal(r["best"], lock["best"])
        assert r["val"] == lock["val"]
        # fewer launches (with lookahead 0 the window follows the measured cost of a call: this objective is cheap, the windows short)
        assert len(r["calls"]) < len(lock["calls"]) / (2 if la else 1)
        assert max(len(c) for c in r["calls"]) <= 4096


def test_many_chains_share_the_launch_budget(host, problem, tmp_path):
    """A launch holds at most 4096 proposals: 1500 chains look 2 ahead, 3000 chains run in lockstep."""
    st = dict(mcmc_iterations=6, burn_in=6, seed=2, write_trace=0, write_checkpoints=0, store_samples=0)
    a = _run(host, problem, tmp_path, "a", "gauss", dict(st, n_chains=1500, lookahead=8))      # 8 asked for, 2 fit (automatic lengths follow measured costs)
    assert max(len(c) for c in a["calls"]) <= 4096 and any(len(c) > 1500 for c in a["calls"])
    b = _run(host, problem, tmp_path, "b", "gauss", dict(st, n_chains=3000))
    assert all(len(c) == 3000 for c in b["calls"][1:]) and b["nev"] == 1 + 3000 * 5
    c = _run(host, problem, tmp_path, "c", "gauss", dict(st, n_chains=1500, lookahead=1))
    np.testing.assert_array_equal(a["best"], c["best"])
    assert a["val"] == c["val"]


def test_window_length_follows_the_cost_structure_of_the_objective(host, problem, tmp_path):
    """The automatic window length comes from measured costs: the first two windows (4 and 16 proposals) split an objective call
    into a fixed and a per-row part.  An objective that scores its rows one after the other gains nothing from looking ahead and
    gets the sequential call pattern back (one row per call); an objective with a fixed cost per call -- a device launch -- gets
    long windows.  The chain is the same chain either way."""
    import time
    inner = _objective(problem, "gauss")
    calls = []

    def serial(x):                       # 0.3 ms per row
        calls.append(len(x))
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < 3e-4 * len(x):
            pass
        return inner(x)

    def fixed(x):                        # 0.6 ms per call
        calls.append(len(x))
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < 6e-4:
            pass
        return inner(x)

    pm = host.ParameterManager(problem.sigmas, problem.lower_bound, problem.upper_bound, mode=1)
    st = dict(mcmc_iterations=400, burn_in=400, n_chains=1, seed=9, write_trace=0, write_checkpoints=0, store_samples=0)
    best = {}
    for name, f in (("serial", serial), ("fixed", fixed)):
        calls.clear()
        best[name] = host.optimize("mh", pm, dict(st, lookahead=0), f, problem.base_params())[:2]
        rows, n_calls = sum(calls), len(calls)
        if name == "serial":
            assert rows < 1.25 * 400, (rows, n_calls)              # the two probe windows, then (nearly) one row per call
        else:
            assert n_calls < 400 / 2 and rows > 2 * 400, (rows, n_calls)
    np.testing.assert_array_equal(best["serial"][0], best["fixed"][0])
    assert best["serial"][1] == best["fixed"][1]
