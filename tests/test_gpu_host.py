"""GPU tests of the C++ host layer (reference-shaped classes over the device C ABI) and of the sharded drivers:
identical accept/reject decisions of a seeded multi-chain Metropolis-Hastings run on the GPU and on the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def host(pkg, cuda_lib):
    import __graft_entry__ as entry
    entry.build()
    from sepaihrd_b200 import hostlib
    return hostlib


@pytest.fixture(scope="module")
def ev_mod(cuda_lib):
    from sepaihrd_b200 import evaluator
    return evaluator


@pytest.fixture(scope="module")
def reflect_problem(problem):
    return problem.__class__.from_json(dict(problem.to_json(), constraint_mode=1))


def _rel(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


def test_objective_function_mirror_matches_oracle(host, problem, oracle):
    """SEPAIHRDObjectiveFunction::calculate / calculateBatch through the C++ mirror (1e-8 relative, north_star)."""
    m = host.HostModel(problem)
    base = problem.base_params()
    np.testing.assert_array_equal(m.current_parameters(), base)                      # getCurrentParameters
    ll_ref = oracle.eval_batch(base[None])[0][0]
    assert _rel(m.calculate(base), ll_ref) < 1e-8
    P = oracle.jitter_params(200, seed=21)
    ref = oracle.eval_batch(P)[0]
    assert _rel(m.calculate_batch(P), ref).max() < 1e-8
    # padded rows (ld > P) and the constraint mode flipped by the parameter manager (clamp <-> reflect)
    wild = base + 50 * problem.sigmas
    clamp_val = m.calculate(wild)
    m.set_constraint_mode(1)
    refl_val = m.calculate(wild)
    rp = problem.__class__.from_json(dict(problem.to_json(), constraint_mode=1))
    import __graft_entry__ as entry
    o_reflect = entry.load_oracle().Oracle(rp)
    assert _rel(clamp_val, oracle.eval_batch(wild[None])[0][0]) < 1e-8
    assert _rel(refl_val, o_reflect.eval_batch(wild[None])[0][0]) < 1e-8
    m.close()


def test_parameter_manager_mirror_updates_the_model(host, problem):
    m = host.HostModel(problem)
    new = np.clip(problem.base_params() * 1.01, problem.lower_bound, problem.upper_bound)
    m.update_parameters(new)                                                        # updateModelParameters -> model
    np.testing.assert_allclose(m.current_parameters(), new, rtol=0, atol=0)
    m.update_parameters(problem.upper_bound + 1.0)                                   # clamped on the way in
    np.testing.assert_array_equal(m.current_parameters(), problem.upper_bound)
    m.close()


def test_simulator_mirror_matches_oracle(host, problem, oracle):
    """AgeSEPAIHRDSimulator::run(initial_state, times): the caller's state integrated as given (1e-6 relative gate)."""
    m = host.HostModel(problem)
    s0 = problem.data_initial_state
    t = problem.times[20:120]
    sub = problem.__class__.from_json(dict(problem.to_json(), times=[float(x) for x in t],
                                           obs_hosp=[float(x) for x in problem.obs_hosp[:100].reshape(-1)],
                                           obs_icu=[float(x) for x in problem.obs_icu[:100].reshape(-1)],
                                           obs_deaths=[float(x) for x in problem.obs_deaths[:100].reshape(-1)]))
    import __graft_entry__ as entry
    ref, st = entry.load_oracle().Oracle(sub).simulate_from_state(problem.base_params()[None], s0)
    got = m.simulate(s0, t)
    assert st[0] == 0
    assert (np.abs(got - ref[0]) / np.maximum(np.abs(ref[0]), 1.0)).max() < 1e-9
    np.testing.assert_array_equal(got[0], s0)
    with pytest.raises(host.HostError):
        m.simulate(s0, t[::-1])                                                      # not strictly increasing
    m.close()


def test_host_mirror_runs_problems_with_other_age_class_counts(host, problem, orc):
    """Three and seven age classes through the C++ mirror (objective, simulator, two-phase calibration, posterior predictive):
    the library pads them to its 4- and 16-lane kernels, the mirror sees the caller's n everywhere."""
    for p in (problem.select_ages([0, 1, 3]), problem.expand_ages(2).select_ages([0, 1, 2, 3, 4, 5, 7])):
        o = orc.Oracle(p)
        m = host.HostModel(p)
        P = o.jitter_params(33, seed=2)
        assert _rel(m.calculate_batch(P), o.eval_batch(P)[0]).max() < 1e-8
        s0 = p.data_initial_state
        ref, st = o.simulate_from_state(p.base_params()[None], s0)
        got = m.simulate(s0, p.times)
        assert got.shape == (p.n_times, 11 * p.n_ages) and (np.abs(got - ref[0]) / np.maximum(np.abs(ref[0]), 1.0)).max() < 1e-9
        f0 = o.eval_batch(p.base_params()[None])[0][0]
        best, val, ns = m.calibrate("hill", dict(iterations=2, cloud_size=16, seed=3), dict(mcmc_iterations=4, burn_in=4, n_chains=4, seed=5))
        assert val >= f0 * (1 - 1e-12) and _rel(o.eval_batch(best[None])[0][0], val) < 1e-8
        q, used = m.posterior_predictive(P[:16], s0)
        assert used == 16 and q.shape[:3] == (6, int((p.times >= 0).sum()), p.n_ages) and np.isfinite(q).all()
        m.close()


def test_seeded_mcmc_accept_sequences_are_identical_on_gpu_and_oracle(problem, reflect_problem, orc, ev_mod, pkg):
    """BASELINE.json configs[2] at test size: every accept/reject decision of a seeded multi-chain run must agree."""
    from sepaihrd_b200 import drivers
    o = orc.Oracle(reflect_problem)
    kw = dict(sigmas=problem.sigmas, lower=problem.lower_bound, upper=problem.upper_bound, initial=problem.base_params(),
              n_chains=96, iterations=25, seed=1234)
    ref = drivers.run_multichain_mh(lambda x: o.eval_batch(x)[0], **kw)
    with ev_mod.BatchEvaluator(reflect_problem, device=0) as ev:
        got = drivers.run_multichain_mh(ev.eval_batch, **kw)
    np.testing.assert_array_equal(got["accepts"], ref["accepts"])
    np.testing.assert_array_equal(got["x"], ref["x"])
    assert _rel(got["logpost"], ref["logpost"]).max() < 1e-8
    assert 0 < ref["accepts"].mean() < 1


def test_seeded_mcmc_4096_chains_with_adaptation_matches_oracle(problem, reflect_problem, orc, ev_mod):
    """BASELINE.json configs[2] at its own width: 4096 seeded chains, with the adaptive part switched on early (burn-in 3,
    covariance recomputed + re-factorised every 2 iterations): identical accept matrix and identical visited states."""
    from sepaihrd_b200 import drivers
    o = orc.Oracle(reflect_problem)
    kw = dict(sigmas=problem.sigmas, lower=problem.lower_bound, upper=problem.upper_bound, initial=problem.base_params(),
              n_chains=4096, iterations=9, seed=1234, settings=dict(burn_in=3, adaptation_period=2))
    ref = drivers.run_multichain_mh(lambda x: o.eval_batch(x)[0], **kw)
    with ev_mod.BatchEvaluator(reflect_problem, device=0) as ev:
        got = drivers.run_multichain_mh(ev.eval_batch, **kw)
    assert got["accepts"].shape == (8, 4096)
    np.testing.assert_array_equal(got["accepts"], ref["accepts"])
    np.testing.assert_array_equal(got["x"], ref["x"])
    np.testing.assert_array_equal(got["scale"], ref["scale"])
    assert _rel(got["logpost"], ref["logpost"]).max() < 1e-8
    assert 0.02 < ref["accepts"].mean() < 0.9


def test_pso_swarm_reaches_the_same_global_best_on_gpu_and_oracle(problem, oracle, ev_mod):
    from sepaihrd_b200 import drivers
    kw = dict(sigmas=problem.sigmas, lower=problem.lower_bound, upper=problem.upper_bound, swarm_size=256, iterations=6, seed=7,
              initial=problem.base_params())
    ref = drivers.run_pso(lambda x: oracle.eval_batch(x)[0], **kw)
    with ev_mod.BatchEvaluator(problem, device=0) as ev:
        got = drivers.run_pso(ev.eval_batch, **kw)
    np.testing.assert_array_equal(got["best_position"], ref["best_position"])
    assert _rel(got["trace"], ref["trace"]).max() < 1e-8
    assert ref["trace"][-1] >= ref["trace"][0]


def test_gradient_objective_is_one_device_batch_of_forward_differences(host, problem, oracle):
    """SEPAIHRDGradientObjectiveFunction::evaluate_with_gradient: (f(x + eps_i e_i) - f(x)) / eps_i, eps_i = 1e-4 max(|x_i|, 1e-4),
    the P perturbed vectors as one device batch, against the same differences of the CPU oracle."""
    m = host.HostModel(problem)
    x = oracle.jitter_params(3, seed=17)[2]
    val, g = m.gradient(x)
    P = len(x)
    steps = 1e-4 * np.maximum(np.abs(x), 1e-4)
    rows = np.tile(x, (P, 1)) + np.diag(steps)
    fc = oracle.eval_batch(x[None])[0][0]
    fp = oracle.eval_batch(rows)[0]
    want = (fp - fc) / steps
    assert _rel(val, fc) < 1e-8
    # the differences amplify the 1e-11 relative agreement of the two evaluators by |f| / (eps_i |df|)
    tol = 1e-7 * np.abs(want) + 4e-10 * abs(fc) / steps
    assert np.all(np.abs(g - want) <= tol), np.max(np.abs(g - want) / tol)
    assert np.count_nonzero(g) >= 50                               # the dead parameters of quirk Q4 have an exactly zero component
    dead = [i for i, nme in enumerate(problem.param_names) if nme.endswith("_multiplier") or nme == "runup_days"]
    assert dead and np.all(g[dead] == 0.0)
    m.close()


def test_nuts_calibration_runs_on_the_device_objective(host, problem, oracle):
    """SEPAIHRDModelCalibration::runNUTS: NUTS as the only phase over the gradient objective (P-vector batches on the device)."""
    m = host.HostModel(problem)
    f0 = oracle.eval_batch(problem.base_params()[None])[0][0]
    best, val, ns = m.calibrate("nuts", {}, dict(nuts_iterations=4, nuts_adaptation_window=2, nuts_max_tree_depth=2, seed=3))
    m.close()
    assert ns == 4 and np.isfinite(val)
    assert _rel(oracle.eval_batch(best[None])[0][0], val) < 1e-8
    assert np.all(best >= problem.lower_bound) and np.all(best <= problem.upper_bound)


def test_objective_consults_the_simulation_cache_like_calculate_does(host, problem, oracle):
    """SimulationCache behind SEPAIHRDObjectiveFunction (ObjectiveFunction.cpp:63-77, 227-234): probe per vector, one device batch
    for the misses, a key repeated inside the batch evaluated once, batches beyond the capacity evaluated whole."""
    m = host.HostModel(problem)
    x = oracle.jitter_params(40, seed=3)
    plain = m.calculate_batch(x)
    assert m.cache_stats() == dict(entries=0, get_calls=0, hits=0, store_calls=0)          # the default is the cache that caches nothing
    m.set_cache(64)
    a = m.calculate(x[0])
    assert m.cache_stats() == dict(entries=1, get_calls=1, hits=0, store_calls=1)
    b = m.calculate(x[0])
    assert a == b == plain[0] and m.cache_stats() == dict(entries=1, get_calls=2, hits=1, store_calls=1)
    rows = [1, 2, 1, 3, 0, 2]
    got = m.calculate_batch(x[rows])
    np.testing.assert_array_equal(got, plain[rows])
    # 6 probes (x0 hits) + 2 re-probes for the repeated keys (hits); 3 distinct misses stored
    assert m.cache_stats() == dict(entries=4, get_calls=2 + 6 + 2, hits=1 + 1 + 2, store_calls=1 + 3)
    np.testing.assert_array_equal(m.calculate_batch(x[:4]), plain[:4])                    # all hits: no device call needed
    assert m.cache_stats()["hits"] == 4 + 4
    m.set_cache(16)                                                                        # 40 sets > 16 entries: evaluated whole
    np.testing.assert_array_equal(m.calculate_batch(x), plain)
    assert m.cache_stats() == dict(entries=0, get_calls=0, hits=0, store_calls=0)
    # least-frequently-used eviction through the objective: 16 entries, then a 17th
    for i in range(17):
        m.calculate(x[i])
    assert m.cache_stats()["entries"] == 16
    m.close()


def test_device_resident_swarm_hands_over_to_the_host_engine_on_a_stagnation_restart(host, problem):
    """restart_threshold = 1e300 makes every iteration "stagnant", so with max_stagnation = 2 the main loop restarts the swarm at
    iteration 3 (ParticleSwarmOptimizer.cpp:133-146).  The device-resident swarm is read back at that point and the run continues
    in the host engine; started on the host it must visit the same states (the device swarm is bit-identical to the host swarm)."""
    s1 = dict(iterations=6, swarm_size=96, seed=3, restart_threshold=1e300, max_stagnation=2, **host.BASIC_SWARM)
    s2 = dict(mcmc_iterations=4, burn_in=4, n_chains=8, seed=5)
    def run(**kw):                      # a calibration leaves its result in the model: every run gets a fresh one
        m = host.HostModel(problem)
        out = m.calibrate("pso", dict(s1, **kw), s2)
        m.close()
        return out
    on_device, on_host, no_restart = run(device_resident=1), run(device_resident=0), run(device_resident=1, max_stagnation=50)
    assert on_device[1] == on_host[1]
    np.testing.assert_array_equal(on_device[0], on_host[0])
    assert not np.array_equal(no_restart[0], on_device[0])        # the restart did change the run


def test_reference_default_swarm_configuration_runs_on_the_device_objective(host, problem, oracle):
    """The shipped pso_settings.txt: von Neumann topology, opposition-based initialisation, adaptive coefficients -- the
    whole-swarm engine with every evaluation as one device batch."""
    m = host.HostModel(problem)
    f0 = oracle.eval_batch(problem.base_params()[None])[0][0]
    s1 = dict(iterations=4, swarm_size=128, seed=3, variant=0, topology=2, use_opposition_learning=1, use_adaptive_parameters=1, max_stagnation=20)
    best, val, ns = m.calibrate("pso", s1, dict(mcmc_iterations=3, burn_in=3, n_chains=8, seed=5))
    m.close()
    assert val >= f0 * (1 - 1e-12)
    assert _rel(oracle.eval_batch(best[None])[0][0], val) < 1e-8


def test_model_calibration_mirror_runs_both_phases(host, problem, oracle):
    """SEPAIHRDModelCalibration::runPSOMCMC / runHillClimbingMCMC end to end on the device (small settings)."""
    m = host.HostModel(problem)
    f0 = oracle.eval_batch(problem.base_params()[None])[0][0]
    best, val, ns = m.calibrate("pso", dict(iterations=3, swarm_size=64, seed=3), dict(mcmc_iterations=12, burn_in=12, n_chains=32, seed=5))
    assert val >= f0 * (1 - 1e-12) and ns == 32 * 12
    assert _rel(oracle.eval_batch(best[None])[0][0], val) < 1e-8
    best, val, ns = m.calibrate("hill", dict(iterations=2, cloud_size=64, seed=3), dict(mcmc_iterations=5, burn_in=5, n_chains=8, seed=5))
    assert val >= f0 * (1 - 1e-12) and ns == 8 * 5
    m.close()


def test_result_aggregator_mirror_matches_the_evaluator(host, problem, oracle, ev_mod):
    """ResultAggregator::aggregatePosteriorPredictives through the C++ mirror == the C-ABI call on the same draws; the
    with-replacement subsampling follows std::mt19937(seed) + uniform_int_distribution."""
    P = oracle.jitter_params(120, seed=8)
    s0 = problem.data_initial_state
    m = host.HostModel(problem)
    got, used = m.posterior_predictive(P, s0)
    with ev_mod.BatchEvaluator(problem, device=0) as ev:
        ref, valid = ev.posterior_predictive(P, s0)
    assert used == valid == 120
    np.testing.assert_array_equal(got, ref)
    sub, used = m.posterior_predictive(P, s0, num_samples=40, seed=99)
    assert used == 40 and sub.shape == got.shape and not np.array_equal(sub, got)
    sub2, _ = m.posterior_predictive(P, s0, num_samples=40, seed=99)
    np.testing.assert_array_equal(sub, sub2)
    m.close()


def _slots_for(problem, vec):
    s = np.array(problem.base_slots, dtype=float)
    for i, sl in enumerate(problem.param_slot):
        if sl >= 0:
            s[sl] = vec[i]
    return s


def test_npi_scenario_analysis_matches_oracle(host, problem, orc, oracle, tmp_path):
    """PostCalibrationAnalyser scenario step (BASELINE.json configs[4]): mean of the kept samples -> baseline, first calibratable
    kappa x0.9 / x1.1, three runs as one device batch.  Trajectories against the CPU oracle (1e-6 gate, observed ~1e-13), metrics
    against the numpy restatement; the scenario runs' metrics model keeps the BASELINE kappa schedule like the reference."""
    from _metrics_ref import essential_metrics, unpack
    d = problem.to_json()
    d["base_slots"] = list(d["base_slots"]); d["base_slots"][problem.layout.beta_scalar] = None      # quirk Q1: no scalar beta
    prob = problem.__class__.from_json(d)
    samples = oracle.jitter_params(40, seed=33)
    burn_in, thinning = 4, 3
    mean = np.clip(samples[burn_in::thinning].sum(axis=0) / len(samples[burn_in::thinning]), prob.lower_bound, prob.upper_bound)
    s0 = oracle.simulate_batch(problem.base_params()[None])[0][0][0]          # the run-up seeded state of calculate()
    m = host.HostModel(prob)
    csv = str(tmp_path / "scenario_comparison.csv")
    out = m.scenarios(samples, s0, burn_in, thinning, csv_path=csv, trajectories=True)
    k2 = prob.param_names.index("kappa_2")
    vecs = np.stack([mean, mean, mean]); vecs[1, k2] *= 0.9; vecs[2, k2] *= 1.1
    lay = prob.layout
    np.testing.assert_allclose(out["kappa"][:, 1], vecs[:, k2], rtol=1e-15)
    np.testing.assert_array_equal(out["kappa"][:, 0], [1.0, 1.0, 1.0])
    # oracle reference with the kappa_2 bounds opened (the scenario values are set on the model directly, unconstrained)
    wide = dict(prob.to_json()); wide["lower_bound"] = list(wide["lower_bound"]); wide["upper_bound"] = list(wide["upper_bound"])
    wide["lower_bound"][k2] = 0.0; wide["upper_bound"][k2] = 10.0
    ref, st = orc.Oracle(prob.__class__.from_json(wide)).simulate_from_state(vecs, s0)
    assert (st == 0).all()
    assert (np.abs(out["trajectories"] - ref) / np.maximum(np.abs(ref), 1.0)).max() < 1e-9
    base_kappa = _slots_for(prob, mean)[lay.kappa0:lay.kappa0 + lay.nk]
    for r in range(3):
        q = unpack(prob, _slots_for(prob, vecs[r]))
        sc, age, _, _ = essential_metrics(q, prob.times, out["trajectories"][r], s0, npi_kappa_values=base_kappa)
        np.testing.assert_allclose(out["scalars"][r], sc, rtol=1e-10)
        np.testing.assert_allclose(out["age"][r], age, rtol=1e-10)
    names = host.METRIC_NAMES
    deaths = out["scalars"][:, names.index("total_cumulative_deaths")]
    assert deaths[1] < deaths[0] < deaths[2]                                  # stricter < baseline < weaker lockdown
    lines = open(csv).read().strip().split("\n")
    assert lines[0].startswith("scenario,R0,overall_IFR,overall_attack_rate,peak_hospital,peak_ICU,") and lines[0].endswith("kappa_7")
    assert [ln.split(",")[0] for ln in lines[1:]] == ["baseline", "stricter_lockdown", "weaker_lockdown"]
    m.close()


def test_mcmc_run_analysis_matches_oracle(host, problem, oracle):
    """analyzeMCMCRunsInBatches: one run per kept sample in one device batch, metrics per run, Rt / seroprevalence quantiles."""
    from _metrics_ref import essential_metrics, unpack
    samples = oracle.jitter_params(30, seed=35)
    s0 = problem.data_initial_state
    m = host.HostModel(problem)
    sc, rt_q, se_q = m.analyze_runs(samples, s0, burn_in=5, thinning=4)
    kept = samples[5::4]
    assert sc.shape == (len(kept), 12)
    ref, st = oracle.simulate_from_state(kept, s0)
    rts, ses = [], []
    for r, vec in enumerate(kept):
        q = unpack(problem, _slots_for(problem, np.clip(vec, problem.lower_bound, problem.upper_bound)))
        s, _, rt, se = essential_metrics(q, problem.times, ref[r], s0)
        np.testing.assert_allclose(sc[r], s, rtol=1e-8)
        rts.append(rt); ses.append(se)
    probs = [0.025, 0.05, 0.5, 0.95, 0.975]

    def quant(cols):
        v = np.sort(np.array(cols), axis=0); n = len(v)
        out = []
        for p in probs:
            pos = p * (n - 1); i = int(pos); f = pos - i
            out.append(v[i] * (1 - f) + v[i + 1] * f if i + 1 < n else v[i])
        return np.array(out)
    np.testing.assert_allclose(rt_q, quant(rts), rtol=1e-8)
    np.testing.assert_allclose(se_q, quant(ses), rtol=1e-8, atol=1e-16)
    m.close()


def test_device_resident_swarm_visits_the_host_swarms_positions(host, problem, ev_mod):
    """sepaihrd_swarm_* (csrc/sepaihrd_swarm.cu): std::mt19937 + uniform_real_distribution and the unfused PSO update on the
    device reproduce the host implementation bit for bit -- same positions, same global-best trace."""
    from sepaihrd_b200 import drivers
    kw = dict(sigmas=problem.sigmas, lower=problem.lower_bound, upper=problem.upper_bound, swarm_size=333, iterations=5, seed=7,
              initial=problem.base_params())
    with ev_mod.BatchEvaluator(problem, device=0) as ev:
        ref = drivers.run_pso(ev.eval_batch, **kw)
        launches0 = ev.counters()[0]
        got = drivers.run_pso(None, device_ctx=ev.handle, **kw)
        launches = ev.counters()[0] - launches0
    np.testing.assert_array_equal(got["trace"], ref["trace"])
    np.testing.assert_array_equal(got["best_position"], ref["best_position"])
    assert got["best_value"] == ref["best_value"] and got["trace"][-1] >= got["trace"][0]
    assert launches == 1 + 6 * 3 + 5          # init, (objective + tell + best) per evaluation, one update per iteration
    # without a starting point every particle is drawn; no-seed runs differ
    kw2 = dict(kw, initial=None, iterations=2)
    with ev_mod.BatchEvaluator(problem, device=0) as ev:
        a = drivers.run_pso(ev.eval_batch, **kw2)
        b = drivers.run_pso(None, device_ctx=ev.handle, **kw2)
    np.testing.assert_array_equal(a["trace"], b["trace"])


def test_device_resident_swarm_shard_matches_host_shard(host, problem, ev_mod):
    """A shard (particle_offset, local_count) of a larger swarm: initial draw, one evaluation, one update towards a given
    global best -- positions identical to the host shard's."""
    pm = host.ParameterManager(problem.sigmas, problem.lower_bound, problem.upper_bound, mode=0)
    st = dict(iterations=9, swarm_size=500, seed=11, particle_offset=137, local_count=91)
    h = host.Swarm(pm, st); d = host.Swarm(pm, st)
    with ev_mod.BatchEvaluator(problem, device=0) as ev:
        h.begin(problem.base_params()); d.begin_device(ev.handle, problem.base_params())
        d.fetch()
        x0 = h.positions()
        np.testing.assert_array_equal(d.positions(), x0)
        assert ((x0 >= problem.lower_bound) & (x0 <= problem.upper_bound)).all() and len(np.unique(x0[:, 0])) == 91
        fit = ev.eval_batch(x0)[0]
        v, i, pos = h.tell(fit)
        dv, di, dpos = d.evaluate_device()
        assert (dv, di) == (v, i)
        np.testing.assert_array_equal(dpos, pos)
        g = np.clip(pos + 0.5 * problem.sigmas, problem.lower_bound, problem.upper_bound)
        h.set_global_best(v + 1.0, g); d.set_global_best(v + 1.0, g)
        for it in (0, 1):
            h.step(it); d.step_device(it)
            d.fetch()
            np.testing.assert_array_equal(d.positions(), h.positions())
            v, i, pos = h.tell(ev.eval_batch(h.positions())[0])
            dv, di, dpos = d.evaluate_device()
            assert (dv, di) == (v, i)
            np.testing.assert_array_equal(dpos, pos)
    # an empty shard and a shard outside the swarm
    with ev_mod.BatchEvaluator(problem, device=0) as ev:
        e = host.Swarm(pm, dict(st, particle_offset=500, local_count=0))
        e.begin_device(ev.handle)
        assert e.evaluate_device()[1] == -1
        with pytest.raises(host.HostError):
            host.Swarm(pm, dict(st, particle_offset=450, local_count=91)).begin_device(ev.handle)


def test_objective_benchmark_harness_runs_from_a_project_tree(host, problem, oracle, tmp_path):
    """host/sepaihrd_objective_benchmark (the reference's timing harness, C++ end to end: tree -> readers -> mirrored objects
    -> device batches): warm-up value and the sum over the harness's own jittered sets against the oracle."""
    import json, os, subprocess
    from sepaihrd_b200 import config
    config.write_reference_tree(problem, str(tmp_path))
    exe = os.path.join(os.path.dirname(host.LIB_PATH), "sepaihrd_objective_benchmark")
    assert os.path.exists(exe), "run `python __graft_entry__.py` (make -C host builds the harness)"
    run = subprocess.run([exe, "--project-root", str(tmp_path), "--repeats", "16", "--jitters", "512", "--seed", "1", "--json"],
                         capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stderr
    rep = json.loads(next(ln for ln in run.stdout.splitlines() if ln.startswith("JSON "))[5:])
    base_ll = oracle.eval_batch(problem.base_params()[None])[0][0]
    jit = oracle.eval_batch(oracle.jitter_params(512, seed=1))[0]
    assert _rel(rep["warmup_value"], base_ll) < 1e-8
    assert _rel(rep["repeat_sum"], 16 * base_ll) < 1e-8
    assert _rel(rep["jitter_sum"], jit.sum()) < 1e-8
    assert "Jitter: 512 evals" in run.stdout and "Objective calls: 529" in run.stdout
    # the batched callers through the same binary: seeded multi-chain MCMC and the device-resident swarm improve on the start
    run = subprocess.run([exe, "--project-root", str(tmp_path), "--mode", "all", "--repeats", "0", "--jitters", "0", "--hill-iters", "3",
                          "--mcmc-iters", "12", "--chains", "64", "--pso-iters", "4", "--swarm", "256", "--json"],
                         capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stderr
    rep = json.loads(next(ln for ln in run.stdout.splitlines() if ln.startswith("JSON "))[5:])
    assert rep["hill_best"] >= base_ll * (1 - 1e-12) and rep["mcmc_best"] >= rep["hill_best"] * (1 - 1e-12) and rep["pso_best"] >= base_ll * (1 - 1e-12)
    # the unchanged reference calling pattern: calculate() from an OpenMP loop with more threads than cores; the C ABI merges
    # concurrent calls into shared launches and every call still gets its own vector's value
    run = subprocess.run([exe, "--project-root", str(tmp_path), "--mode", "threads", "--threads", "128", "--jitters", "2048", "--seed", "1", "--json"],
                         capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stderr
    rep = json.loads(next(ln for ln in run.stdout.splitlines() if ln.startswith("JSON "))[5:])
    assert _rel(rep["threads_sum"], oracle.eval_batch(oracle.jitter_params(2048, seed=1))[0].sum()) < 1e-8
    assert rep["threads_launches"] < 2048 / 4
    bad = subprocess.run([exe, "--project-root", str(tmp_path / "nowhere")], capture_output=True, text=True)
    assert bad.returncode == 1 and "unable to open" in bad.stderr


def test_single_chain_lookahead_is_the_sequential_chain_on_the_device_objective(host, problem, oracle):
    """The reference's shipped phase 2 -- one chain (MetropolisHastingsSampler.cpp:283-384) -- on the device objective: with look-ahead
    (the next K iterations' proposals in ONE launch, host/optimizers.cpp runLookahead) the chain ends in the state, with the
    log-posterior, scale, acceptance rate and best value of the one-evaluation-per-launch run, in a fraction of the launches;
    through the burn-in, rank-1 covariance updates and two refactorisations of the proposal kernel."""
    m = host.HostModel(problem)
    x0 = problem.base_params()
    st = dict(mcmc_iterations=260, burn_in=100, adaptation_period=60, n_chains=1, seed=8, store_samples=0, write_trace=0, write_checkpoints=0)
    seq = m.metropolis(dict(st, lookahead=1), x0)
    assert seq["launches"] == 260 and seq["evaluations"] == 260 and seq["iterations"] == 260
    for la in (0, 5, 48):
        r = m.metropolis(dict(st, lookahead=la), x0)
        np.testing.assert_array_equal(r["last"], seq["last"])
        assert (r["last_logpost"], r["best_value"], r["final_scale"], r["acceptance_rate"]) == \
               (seq["last_logpost"], seq["best_value"], seq["final_scale"], seq["acceptance_rate"])
        np.testing.assert_array_equal(r["best"], seq["best"])
        assert r["launches"] < seq["launches"] and r["evaluations"] > seq["evaluations"]
    assert _rel(oracle.eval_batch(seq["last"][None])[0][0], seq["last_logpost"]) < 1e-8
    m.close()
